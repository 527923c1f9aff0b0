// forwards to the oracle-side cv:: stand-in (test infrastructure only)
#include "minicv.hpp"
