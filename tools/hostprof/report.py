"""Summary of a PC sample file written by pcsample.c: share per module, function and source line of one library (built with -g).
usage: report.py samples.pcs /path/to/libarriba_b200_hostsim_g.so"""
import collections, subprocess, sys


def main():
    path, lib = sys.argv[1], sys.argv[2]
    mods = collections.Counter(); offs = collections.Counter()
    for l in open(path):
        m, off, _ = l.rstrip("\n").split("\t")
        mods[m.split("/")[-1]] += 1
        if m == lib or m.split("/")[-1] == lib.split("/")[-1]:
            offs[off] += 1
    total = sum(mods.values())
    print("samples", total, dict(mods.most_common(6)))
    addrs = list(offs)
    out = subprocess.run(["addr2line", "-f", "-C", "-e", lib] + ["0x" + a for a in addrs], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    fn = collections.Counter(); ln = collections.Counter()
    for i, a in enumerate(addrs):
        fn[out[2 * i][:100]] += offs[a]
        ln[out[2 * i + 1].split("/")[-1].split(" ")[0]] += offs[a]
    print("-- functions")
    for f, c in fn.most_common(30):
        print("%6.1f%%  %s" % (100.0 * c / total, f))
    print("-- lines")
    for f, c in ln.most_common(50):
        print("%6.1f%%  %s" % (100.0 * c / total, f))


if __name__ == "__main__":
    main()
