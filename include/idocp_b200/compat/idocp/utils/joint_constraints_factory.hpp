// forwarding header: idocp/utils/joint_constraints_factory.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
