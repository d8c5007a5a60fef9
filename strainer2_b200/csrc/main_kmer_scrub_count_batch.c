/* drop-in executable: everything lives in libstrainer2_b200.so.  The process ends with _exit() after flushing its
 * streams: tearing the CUDA context down at exit costs up to seconds and frees nothing the OS does not free anyway. */
#include <stdio.h>
#include <unistd.h>
int s2_kmer_scrub_count_batch_main(int argc, char **argv);
int main(int argc, char **argv)
{
    const int rc = s2_kmer_scrub_count_batch_main(argc, argv);
    fflush(NULL);
    _exit(rc);
}
