// forwarding header: idocp/hybrid/discrete_event.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
