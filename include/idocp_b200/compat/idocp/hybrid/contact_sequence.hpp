// forwarding header: idocp/hybrid/contact_sequence.hpp -> idocp_b200 (see ../../idocp_b200_compat.hpp)
#include "../../idocp_b200_compat.hpp"
