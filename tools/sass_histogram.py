"""SASS opcode histogram of the built library (not a test): python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.json
One entry per kernel: instruction count and the counts of the opcodes that show what the kernel is made of (TMA bulk
copies, mbarriers, packed FP32, FP64, MUFU, tensor-core ops -- there are none, by design -- ...)."""
import collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "skeleton_action_recognition_b200", "lib", "libvirtual_radar_b200.so")
KEYS = ["UBLKCP", "UBLKPF", "UBLKRED", "UTMALDG", "UTMASTG", "SYNCS", "BAR", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FMNMX", "FMNMX3",
        "MUFU", "DFMA", "DADD", "DMUL", "F2F", "LDS", "STS", "LDG", "STG", "LDC", "LDCU", "ATOMG", "RED", "SHFL", "NANOSLEEP", "ACQBULK",
        "UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "HMMA", "IMMA", "USETMAXREG", "ERRBAR", "MEMBAR"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
out, cur, arch = {}, None, set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = out.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur is not None:
        cur["_total"] += 1
        cur[m.group(1)] += 1
        if m.group(1) == "MUFU" and m.group(2):
            cur["MUFU." + m.group(2).split(".")[1]] += 1
res = {"arch": sorted(arch), "kernels": {}}
for name, c in out.items():
    d = {"instructions": c["_total"]}
    d.update({k: c[k] for k in KEYS if c[k]})
    d.update({k: v for k, v in c.items() if k.startswith("MUFU.")})
    res["kernels"][name] = d
json.dump(res, sys.stdout, indent=1)
