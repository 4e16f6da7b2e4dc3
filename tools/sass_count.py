"""Static view of the fast kernel's SASS: registers/spills and the instruction count of the hot record loop
(between the 128-bit pre-state load and the backward branch).  usage: python tools/sass_count.py [extra nvcc flags]"""
import re, subprocess, sys, collections
src = "abm_amro_b200/csrc/kernels_fast.cu"
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false", "-Xptxas", "-v",
       "-cubin", "-o", "/tmp/kf.cubin", src] + sys.argv[1:]
out = subprocess.run(cmd, capture_output=True, text=True)
for line in out.stderr.splitlines():
    if "error" in line: print(line)
    if "k_fastILb0" in line:
        grab = 3
    if "registers" in line or "spill" in line:
        print(line.strip())
sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN4amro6k_fastILb0ELi128ELi4EEEvNS_8FastArgsE", "/tmp/kf.cubin"], capture_output=True, text=True).stdout
ins = []
for l in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print("total SASS instructions:", len(ins))
# loops: backward branches
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`?\(?\.?L_x_\d+\)?|BRA.*0x([0-9a-f]+)", t)
    m2 = re.search(r"0x([0-9a-f]+)", t) if t.split()[0].lstrip("@!UP0123456789 ").startswith("BRA") or " BRA" in t else None
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt in addr_index and tgt < a:
            loops.append((addr_index[tgt], i))
for s, e in loops:
    body = ins[s:e + 1]
    ops = collections.Counter((t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0] for _, t in body)
    has_ldg128 = any("LDG.E.128" in t for _, t in body)
    if has_ldg128:
        print(f"loop [{s},{e}] len {e - s + 1}  (per update if straight-line: {(e - s + 1) / 8:.1f})", dict(ops.most_common(12)))
