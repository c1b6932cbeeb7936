// Library identity.
#include "../../include/gradflow_b200.h"

extern "C" int gf_version(void) { return 100; }
