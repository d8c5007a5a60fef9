"""TEST INFRASTRUCTURE - CPU restatement of the reference's kmer_scrub_filter step.

Follows /root/reference/scripts/kmer_scrub_filter.py: table reading and hash merging :153-202, headers :204-216,
drug scrub :62-68, independent scrub :31-58 + :72-84, joint scrub :88-143, output :226-229.  Parity is PINNED:
tests/test_filter.py checks this restatement against tests/golden/cases/filter/, which
tests/golden/make_golden_filter.py produced by running the unmodified script.  Only tests/ may import this file;
the product (strainer2_b200/bin/kmer_scrub_filter) never does.

run(argv, cwd) -> (exit code, stdout bytes, stderr bytes); a stderr that starts with "<traceback>" stands for an
uncaught Python exception whose last line follows."""
import gzip
import os


class _Exit(Exception):
    def __init__(self, code, msg):
        self.code, self.msg = code, msg


def _parse_args(argv):
    """the four options of the script's argparse parser (:14-27); enough of argparse for the tests"""
    opt = {"scrub_count_file": None, "scrub_count_list": None, "min_fraction": 0.04, "independent": False}
    names = {"-s": "scrub_count_file", "--scrub_count_file": "scrub_count_file", "-l": "scrub_count_list",
             "--scrub_count_list": "scrub_count_list", "-m": "min_fraction", "--min_fraction": "min_fraction"}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-i", "--independent"):
            opt["independent"] = True
        elif a in names:
            i += 1
            v = argv[i]
            opt[names[a]] = float(v) if names[a] == "min_fraction" else v
        else:
            raise _Exit(2, "unrecognized arguments: " + a)
        i += 1
    return opt


def _threshold_scrub(min_frac, counts, total, err):
    """scrub_max_kmers :31-58: raise the threshold until at least min_frac of the k-mers would be kept"""
    t, kept = -1, -1.0
    total = float(total)
    while kept < min_frac:
        t += 1
        hits = sum(1 for v in counts.values() if v > t)
        kept = 1 - (hits / total)
        err.append("kept " + str(kept) + " with threshold " + str(t) + "\n")
    scrub = {k: v for k, v in counts.items() if v > t}
    err.append("threshold was " + str(t) + " left with " + str(len(scrub)) + " out of " + str(total) + " that will be scrubbed\n")
    return scrub


def run(argv, cwd="."):
    out, err = [], []
    try:
        opt = _parse_args(argv)
        m = opt["min_fraction"]
        if m < 0.0 or m > 1.0:                                   # :147-148: str + float raises
            raise _Exit(1, "<traceback>\nTypeError: can only concatenate str (not \"float\") to str\n")
        if not opt["scrub_count_file"] and not opt["scrub_count_list"]:
            err.append("error: one of scrub_count_file or scrub_count_list must be provided.")
        if opt["scrub_count_file"] and opt["scrub_count_list"]:
            err.append("error: can provide only one of either scrub_count_file or scrub_count_list.")
        files = []
        if opt["scrub_count_file"]:
            files.append(opt["scrub_count_file"])
        elif opt["scrub_count_list"]:
            files = [line.rstrip() for line in open(os.path.join(cwd, opt["scrub_count_list"]))]

        strain, meta, pan, drug = {}, {}, {}, {}
        drug_filter, all_kmers = 0, 0
        previous = None
        for i, f in enumerate(files):
            if i > 1:                                            # :163 (sic): only from the third file on
                previous = strain
            strain, all_kmers = {}, 0
            with gzip.open(os.path.join(cwd, f), "rt") as reader:
                for line in reader:
                    if line.startswith("#"):
                        continue
                    c = line.rstrip("\n").split("\t")
                    all_kmers += 1
                    strain[c[0]] = int(c[1])
                    if int(c[2]) > 0:
                        pan[c[0]] = pan.get(c[0], 0) + int(c[2])
                    if int(c[3]) > 0:
                        meta[c[0]] = meta.get(c[0], 0) + int(c[3])
                    if len(c) == 5:
                        drug_filter = 1
                        if int(c[4]) > 0:
                            drug[c[0]] = drug.get(c[0], 0) + int(c[3])       # :194 (sic): membership is what matters
            if i > 1 and strain != previous:
                raise _Exit(1, "error: input files do not have identical hash and strain hash values.\n")

        out.append("#total kmers in strain:%d,%d pangenome: %d metagenome: %d\n" % (all_kmers, len(strain), len(pan), len(meta)))
        drug_scrubbed = 0
        if drug_filter:
            out.append("#total kmers cross drug:%d\n" % len(drug))
            for k in drug:
                strain.pop(k, None)
            remaining = float(len(strain) / float(all_kmers))
            drug_scrubbed = all_kmers - len(strain)
            out.append("#fraction kmers remaining drug post scrub:" + str(remaining) + "\n")
            out.append("#drug_scrubbed kmers:" + str(drug_scrubbed) + "\n")
            if remaining < m * 2:
                raise _Exit(1, "<traceback>\nException: ERROR: too few kmers remain after drug scrub. Are your drug strains too similar?\n")

        if opt["independent"]:
            for scrub in (_threshold_scrub(m, pan, all_kmers, err), _threshold_scrub(m, meta, all_kmers, err)):
                for k in scrub:
                    strain.pop(k, None)
        else:
            msum = sum(meta.values())
            meta = {k: v / float(msum) for k, v in meta.items()}
            psum = sum(pan.values())
            pan = {k: v / float(psum) for k, v in pan.items()}
            value = {}
            for k in strain:                                     # the larger of the two fractions, 0 when in neither
                v = 0
                if k in meta and meta[k] > v:
                    v = meta[k]
                if k in pan and pan[k] > v:
                    v = pan[k]
                value[k] = v
            order = sorted(value.items(), key=lambda kv: kv[1], reverse=True)    # stable: ties keep row order
            n = float(drug_scrubbed)
            for k, _ in order:
                if (1 - ((n + 1) / all_kmers)) > m:
                    n += 1.0
                    del strain[k]
        out.append("#post scrub kmers %d out of %d\n" % (len(strain), all_kmers))
        out.extend(k + "\n" for k in strain)
        return 0, "".join(out).encode(), "".join(err).encode()
    except _Exit as e:
        return e.code, "".join(out).encode(), ("".join(err) + e.msg).encode()
