"""Per-kernel counts of the SASS mnemonics that show a Blackwell-native kernel (B200_PROFILING.md: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA; HMMA would be the legacy mma.sync path) in the built objects.
usage: python tools/sass_evidence.py > profiles/sass_evidence.txt   (no GPU needed: cuobjdump on csrc/*.o)"""
import glob, os, re, subprocess, collections
HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "myrtle-vision_b200", "csrc")
PAT = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
       ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("REDG", r"\bREDG\."),
       ("SHFL", r"\bSHFL"), ("MUFU", r"\bMUFU")]
print("%-78s %s" % ("kernel (demangled, truncated)", "  ".join("%8s" % n for n, _ in PAT)))
for obj in sorted(glob.glob(os.path.join(HERE, "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, counts = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = [0] * len(PAT)
            continue
        if name:
            for i, (_, p) in enumerate(PAT):
                if re.search(p, line):
                    counts[name][i] += 1
    if not counts:
        continue
    names = list(counts)
    dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print("---- " + os.path.basename(obj))
    for n, d in zip(names, dem if len(dem) == len(names) else names):
        if sum(counts[n][:6]) == 0 and "gemm" not in d and "attn" not in d:
            continue                      # elementwise kernels: nothing to show
        print("%-78s %s" % (d[:78], "  ".join("%8d" % c for c in counts[n])))
