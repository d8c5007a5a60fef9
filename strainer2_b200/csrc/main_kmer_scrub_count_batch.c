/* kmer_scrub_count_batch - many strains, one pass over the scrub lists (no reference counterpart). */
#include "../../include/strainer2_b200.h"
int main(int argc, char **argv) { return s2_kmer_scrub_count_batch_main(argc, argv); }
