// forwarding header: idocp/cost/contact_force_cost.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
