/* SwiftPM needs one translation unit per C target; the code lives in libflux2b.so (linked via the module map). */
#include "flux2b.h"
