// oracle/ref_shim: TEST INFRASTRUCTURE (see lsd_ref_shim_cv.hpp)
#include "../lsd_ref_shim_cv.hpp"
