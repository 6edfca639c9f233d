// forwarding header: idocp/constraints/impulse_friction_cone.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
