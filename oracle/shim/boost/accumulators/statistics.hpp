// Oracle build shim: see accumulators.hpp.
#include "accumulators.hpp"
