/* strain_detect - drop-in executable (replaces /root/reference/src/strain_detect.c). */
#include "../../include/strainer2_b200.h"
int main(int argc, char **argv) { return s2_strain_detect_main(argc, argv); }
