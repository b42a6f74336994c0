// oracle/ref_shim: TEST INFRASTRUCTURE (see lsd_ref_shim_boost.hpp)
#include "../lsd_ref_shim_boost.hpp"
