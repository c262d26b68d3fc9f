"""Line-level instruction counts of one kernel from an `ncu --set full --import-source on` capture.
The capture's source page lists SASS only; this joins it, instruction by instruction, with the line table of the
same kernel in the built library (nvdisasm -g), so the library must be the build the capture was taken from.
    python tools/ncu_lines.py gpurun_out/r02_prof_x.ncu-rep k_bor_pixel_packed [top]
Prints, per source line: share of executed warp instructions, share of stall samples."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("DOFS3D_LIB", os.path.join(ROOT, "denseopticalflowsegmentation3d_b200", "libdofs3d.so"))


def sass_lines(kernel):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout
    lines = text.split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l]
    if not starts:
        raise SystemExit(f"no kernel matching {kernel!r} in {LIB}")
    out, cur = [], None
    for l in lines[starts[0] + 1:]:
        if l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            out.append(cur)
    return out


def main(rep, kernel, top=30):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H = rows[1]
    ci, si = H.index("Instructions Executed"), H.index("# Samples")
    data = [(float(r[ci]), float(r[si])) for r in rows[2:] if len(r) > ci and r[ci].replace(".", "").isdigit()]
    loc = sass_lines(kernel)
    if len(loc) != len(data):
        raise SystemExit(f"{len(data)} instructions in the capture, {len(loc)} in the library: not the same build")
    inst, samp = collections.Counter(), collections.Counter()
    for l, (v, s) in zip(loc, data):
        inst[l] += v
        samp[l] += s
    ti, ts = sum(inst.values()), max(1.0, sum(samp.values()))
    print(f"{rows[0][1][:70]}: {ti:.0f} warp instructions, {len(data)} SASS instructions")
    cache = {}
    for l, v in inst.most_common(top):
        txt = ""
        if l:
            for d in ("denseopticalflowsegmentation3d_b200/csrc",):
                path = os.path.join(ROOT, d, l[0])
                if os.path.exists(path):
                    cache.setdefault(path, open(path).read().split("\n"))
                    txt = cache[path][l[1] - 1].strip()[:96]
        print(f"{100 * v / ti:5.1f}% inst {100 * samp[l] / ts:5.1f}% stall  {l[0] if l else '?'}:{l[1] if l else 0}  {txt}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
