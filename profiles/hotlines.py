#!/usr/bin/env python3
"""Per-source-line stall samples / instruction counts of one kernel of an .ncu-rep (captured with --import-source on
and a -lineinfo build).  ncu's CSV source page is per SASS instruction; nvdisasm -g on the cubin of the SAME
librdcgpu.so gives the line of every instruction, and the two lists have the same order.
    python profiles/hotlines.py gpurun_out/prof_asm_r1c.ncu-rep k_assemble assemble k_assembleINS_4AdpmELi4ELi128ELi4
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kregex, cubin_stem, mangled = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "rdcfes_b200", "librdcgpu.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# one section per profiled launch: "Kernel Name",<name> / header / one row per SASS instruction
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sections = [(rows[a][1], rows[a + 1], rows[a + 2:b]) for a, b in zip(starts[:-1], starts[1:])]
want = sys.argv[5] if len(sys.argv) > 5 else ""
name, hdr, data = [sec for sec in sections if want in sec[0]][0]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.startswith(cubin_stem + ".") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l][0]
cur, seq = None, []
for l in dis[start + 1:]:
    if (l.startswith(".text.") or l.startswith(".section")) and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
for key, r in zip(seq, data):
    agg[key][0] += int(r[si]); agg[key][1] += int(r[ii])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
src = {f: open(os.path.join(root, "rdcfes_b200", "csrc", f)).read().splitlines() for f in os.listdir(os.path.join(root, "rdcfes_b200", "csrc"))}
print(f"# {rep} kernel {name[:80]}: {tot} stall samples, {toti} warp instructions")
print("# share of samples | share of instructions | file:line | source")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    f, l = key if key else ("?", 0)
    text = src[f][l - 1].strip()[:100] if f in src and 0 < l <= len(src[f]) else ""
    print(f"{100*v[0]/tot:5.1f}% {100*v[1]/toti:5.1f}%  {f}:{l:<4} {text}")
