// forwarding header: idocp/cost/task_space_6d_cost.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
