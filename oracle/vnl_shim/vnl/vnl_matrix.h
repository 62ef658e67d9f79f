// TEST INFRASTRUCTURE ONLY: VNL stand-in header (see vnl_shim_core.h).
#include "vnl_shim_core.h"
