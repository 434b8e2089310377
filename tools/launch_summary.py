"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals/shares and
the kernel sequence of the last MSM step.  usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("b200zk::", "").replace("void ", "")
        t = float(row["Metric Value"].replace(",", ""))
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        seq.append((int(row["ID"]), name, t))
    return seq


def main(path):
    seq = load(path)
    agg = collections.OrderedDict()
    for _, n, t in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    print("%-44s %5s %12s %7s" % ("kernel", "n", "total ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %5d %12.3f %6.1f%%" % (k[:44], v[0], v[1], 100 * v[1] / tot))
    idx = [i for i, s in enumerate(seq) if "msm_accumulate" in s[1]]
    if idx:
        i = idx[-1]
        j = i
        while j > 0 and "msm_digits_kernel<0>" not in seq[j][1]:
            j -= 1
        while j > 0 and ("msm_coarse" in seq[j - 1][1] or "msm_fine" in seq[j - 1][1] or
                         ("scan_tile" in seq[j - 1][1] and j > 1 and "msm_coarse" in seq[j - 2][1])):
            j -= 1
        k = i
        while k < len(seq) and "msm_combine" not in seq[k][1]:
            k += 1
        step = seq[j:k + 1]
        st = sum(s[2] for s in step)
        print("\nlast MSM step (%d launches, %.3f ms serialised):" % (len(step), st))
        for s in step:
            print("  %-42s %10.4f ms %5.1f%%" % (s[1][:42], s[2], 100 * s[2] / st))


if __name__ == "__main__":
    main(sys.argv[1])
