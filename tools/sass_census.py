"""Mnemonic census of the production kernels in libparesis_b200.so (cuobjdump -sass): what the SASS uses and what it does not.
python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paresis_b200", "libparesis_b200.so")
KERNELS = [("object hop", "refract_lean_kernelILi2ELb1ELb1ELi16ELi256ELb0"), ("membrane hop", "refract_lean_kernelILi1ELb0ELb0ELi16ELi256ELb0"),
           ("detector", "detect_tile_kernelILi2ELi1ELi4"), ("membrane cut", "membrane_from_field_batch_kernelILi4"),
           ("splat strips (2 columns)", "splat_strip2_kernel"), ("Fresnel line transform M=8192", "line_convolve_kernelILi13"),
           ("Fresnel margin terms", "post_lines_kernelILi15")]
WATCH = ["ATOMS", "RED", "ATOMG", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "STG.E.64", "LDS.128", "LDS.64", "STS.128", "STS.64", "BAR.SYNC",
         "BAR.ARV", "IMAD.HI", "IMAD.WIDE", "MUFU", "SHFL", "VOTE", "FFMA", "LDL", "STL", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "UTCMMA"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
print("cuobjdump -sass %s   cubin architectures in the library: %s" % (os.path.basename(LIB), sorted(set(re.findall(r"arch = (sm_\w+)", txt)))))
for label, key in KERNELS:
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if key not in name:
            continue
        ops = [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", b)]
        c = collections.Counter()
        for o in ops:
            for w in WATCH:
                if o == w or o.startswith(w + ".") or (w.count(".") and o.startswith(w)):
                    c[w] += 1
        # the generic LDG.E row counts every width; subtract the wide ones for the 32-bit figure
        c["LDG.E (32-bit)"] = c.pop("LDG.E", 0) - c["LDG.E.128"] - c["LDG.E.64"]
        print("\n%s  --  %s\n  static instructions: %d" % (label, name[:110], len(ops)))
        print("  " + ", ".join("%s %d" % (k, c[k]) for k in sorted(c) if c[k]))
        print("  absent: " + ", ".join(w for w in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "UTCMMA", "LDL", "STL") if not c[w]))
        break
    else:
        print("\n%s: no function matching %s" % (label, key))
