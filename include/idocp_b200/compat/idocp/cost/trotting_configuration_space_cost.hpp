// forwarding header: idocp/cost/trotting_configuration_space_cost.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
