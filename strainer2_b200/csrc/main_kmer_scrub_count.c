/* kmer_scrub_count - drop-in executable (replaces /root/reference/src/kmer_scrub_count.c). */
#include "../../include/strainer2_b200.h"
int main(int argc, char **argv) { return s2_kmer_scrub_count_main(argc, argv); }
