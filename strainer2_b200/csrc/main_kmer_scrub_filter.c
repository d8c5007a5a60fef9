/* drop-in for scripts/kmer_scrub_filter.py: everything lives in libstrainer2_b200.so */
int s2_kmer_scrub_filter_main(int argc, char **argv);
int main(int argc, char **argv) { return s2_kmer_scrub_filter_main(argc, argv); }
