/* TEST INFRASTRUCTURE ONLY.
 * Drives the reference's vendored FASTA/FASTQ parser (kseq.h, included from /root/reference/src
 * via -I, never copied) and prints, for every kseq_read() call, "<ret>\t<seq.l>\t<seq>\n" so
 * that our own reader (strainer2_b200/csrc/s2_reader.cpp) can be pinned against it, including
 * the terminating negative return code (-1 EOF / -2 truncated quality).
 * Mirrors the call pattern of src/genome_compare.c:194-203. */
#include <zlib.h>
#include <stdio.h>
#include "kseq.h"
KSEQ_INIT(gzFile, gzread)

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: kseq_dump <file>\n"); return 2; }
    gzFile fp = gzopen(argv[1], "r");
    if (!fp) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    kseq_t *seq = kseq_init(fp);
    int l;
    while ((l = kseq_read(seq)) >= 0) {
        printf("%d\t%zu\t", l, seq->seq.l);
        fwrite(seq->seq.s, 1, seq->seq.l, stdout);
        putchar('\n');
    }
    printf("%d\t%zu\t<END>\n", l, seq->seq.l);
    kseq_destroy(seq);
    gzclose(fp);
    return 0;
}
