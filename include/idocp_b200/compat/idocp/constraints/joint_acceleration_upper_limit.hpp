// forwarding header: idocp/constraints/joint_acceleration_upper_limit.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
