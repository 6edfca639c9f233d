// forwarding header: idocp/cost/time_varying_task_space_3d_cost.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
