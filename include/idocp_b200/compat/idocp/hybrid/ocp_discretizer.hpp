// forwarding header: idocp/hybrid/ocp_discretizer.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
