"""SASS evidence per kernel of dl-unet_b200/csrc/libunet_b200.so: counts of the Blackwell-native instructions
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld; ".2CTA" = cta_group::2 forms of the MMA,
the TMA loads and the multicast commit; LDGMC = multimem.ld_reduce through the NVSwitch multicast address) and of floating-point global atomics
(RED / ATOM .ADD on fp32: none may remain in the training path -- deterministic gradients; ATOMG.EXCH / integer atomics are
the dead-lock guard of mbar_wait and the loss kernel's block counter).   python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "dl-unet_b200", "csrc", "libunet_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
counts = collections.OrderedDict()
fn, k = None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        fn = re.sub(r"\(.*", "", names[k]).replace("void ", "").replace("ub::", "")
        counts[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    for key, pat in (("UTCHMMA", r"\bUTCHMMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"),
                     ("HMMA", r"\bHMMA\."), ("2CTA", r"\.2CTA"), ("LDGMC", r"\bLDGMC"), ("fp32_atomic_add", r"\b(RED|ATOMG|ATOM)\.E\.ADD\.F32|\bRED\.E\.ADD\.F32|ATOMG\.E\.ADD\.F32|\bRED\.E\.ADD\.F32x"),
                     ("other_atomics", r"\b(ATOMG|ATOM|RED)\.")):
        if re.search(pat, line):
            counts[fn][key] += 1
print(f"{'kernel':58s} {'UTCHMMA':>8s} {'UTMALDG':>8s} {'UTMASTG':>8s} {'LDTM':>6s} {'.2CTA':>6s} {'LDGMC':>6s} {'HMMA':>5s} {'f32 atomic add':>15s} {'other atomics':>14s}")
for fn, c in counts.items():
    print(f"{fn[:58]:58s} {c['UTCHMMA']:8d} {c['UTMALDG']:8d} {c['UTMASTG']:8d} {c['LDTM']:6d} {c['2CTA']:6d} {c['LDGMC']:6d} {c['HMMA']:5d} {c['fp32_atomic_add']:15d} "
          f"{c['other_atomics'] - c['fp32_atomic_add']:14d}")
