"""Per-kernel SASS opcode summary of the built library + the full listing of the hot kernel, for profiles/.
usage: python tools/sass_summary.py [round-tag, default r02]"""
import collections, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = "adcraft_b200/_build/libadcraft_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input=txt, capture_output=True, text=True).stdout or txt
kernels, cur, name = collections.OrderedDict(), None, None
for line in filt.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        name = m.group(1).strip()
        cur = kernels.setdefault(name, [])
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        cur.append(m.group(1))
WATCH = ("IMAD.WIDE.U32", "LOP3.LUT", "POPC", "REDUX", "ATOMS", "ATOMG", "UBLKCP", "SYNCS", "LDG", "STG", "DFMA", "SHFL", "VOTE", "MATCH")
out = [f"SASS opcode summary of {lib} (sm_100a), round {tag[1:].lstrip('0')}",
       f"produced by: python tools/sass_summary.py (cuobjdump -sass <lib>) ; full listing of the hot kernel: {tag}_sass_adc_flat2_implicit_kernel.txt", ""]
for k, ops in kernels.items():
    c = collections.Counter(ops)
    fam = {w: sum(v for o, v in c.items() if o == w or o.startswith(w + ".") or (w in ("LDG", "STG", "REDUX", "ATOMS", "ATOMG", "SHFL", "VOTE", "SYNCS", "UBLKCP", "MATCH") and o.startswith(w))) for w in WATCH}
    out += [k, f"  instructions: {len(ops)}", "  " + ", ".join(f"{w}: {n}" for w, n in fam.items() if n),
            "  top: " + ", ".join(f"{o} {n}" for o, n in c.most_common(12)), ""]
open(f"profiles/{tag}_sass_summary.txt", "w").write("\n".join(out))
# full listing of the hot kernel (both instantiations)
keep, on = [], False
for line in filt.splitlines():
    if re.match(r"\s*Function : ", line):
        on = "adc_flat2_implicit_kernel" in line
    if on or line.startswith("Fatbin") or line.startswith("arch =") or line.startswith("code version"):
        keep.append(line)
open(f"profiles/{tag}_sass_adc_flat2_implicit_kernel.txt", "w").write("\n".join(keep) + "\n")
print(len(kernels), "kernels;", sum(len(v) for v in kernels.values()), "instructions")
