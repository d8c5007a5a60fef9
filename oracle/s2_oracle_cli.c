/* s2_oracle_cli.c - TEST INFRASTRUCTURE ONLY: command-line front end of the CPU oracle.
 *   oracle_cli count  -r R -A listA -B listB [-C listC] [-p progress]      (table on stdout)
 *   oracle_cli detect -r R -a informative (-B batch | -b f1 [-c f2] [-t T]) (hit text on stdout,
 *                                                          reference-stdout messages on stderr fd 3 -> see -m)
 *   oracle_cli kseq FILE      (same dump format as oracle/_ref/kseq_dump)
 */
#include "s2_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: oracle_cli count|detect|kseq ...\n"); return 2; }
    const char *cmd = argv[1];
    if (!strcmp(cmd, "kseq")) {
        s2o_reader *r = s2o_reader_open(argv[2]);
        int l;
        if (!r) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
        while ((l = s2o_reader_next(r)) >= 0) {
            printf("%d\t%zu\t", l, s2o_reader_len(r));
            fwrite(s2o_reader_seq(r), 1, s2o_reader_len(r), stdout);
            putchar('\n');
        }
        printf("%d\t%zu\t<END>\n", l, s2o_reader_len(r));
        s2o_reader_close(r);
        return 0;
    }
    const char *r = NULL, *A = NULL, *B = NULL, *C = NULL, *p = NULL, *a = NULL, *b = NULL, *c = NULL, *t = NULL, *m = NULL, *g = NULL;
    int o;
    optind = 2;
    while ((o = getopt(argc, argv, "r:A:B:C:p:a:b:c:t:m:g:")) != -1)
        switch (o) {
        case 'r': r = optarg; break; case 'A': A = optarg; break; case 'B': B = optarg; break;
        case 'C': C = optarg; break; case 'p': p = optarg; break; case 'a': a = optarg; break;
        case 'b': b = optarg; break; case 'c': c = optarg; break; case 't': t = optarg; break;
        case 'm': m = optarg; break;
        case 'g': g = optarg; break;
        default: return 2;
        }
    if (!strcmp(cmd, "count")) return s2o_kmer_scrub_count(r, A, B, C, p, stdout, stderr);
    if (!strcmp(cmd, "detect")) {
        FILE *msg = m ? fopen(m, "w") : stderr;   /* -m FILE: where the reference's stdout chatter goes */
        int rc = s2o_strain_detect_g(r, a, g, B, b, c, t, stdout, msg, stderr);
        if (m) fclose(msg);
        return rc;
    }
    fprintf(stderr, "unknown command %s\n", cmd);
    return 2;
}
