// placeholder, replaced below
#include "../../include/strainer2_b200.h"
extern "C" int s2_strain_detect_main(int, char **) { return 1; }
