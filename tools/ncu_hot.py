"""Hot spots of an ncu capture (--set full --import-source on), per SASS instruction and per source line.

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def load(rep, mode):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", mode],
                         capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"Line No"') or l.startswith('"#"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = load(rep, "sass")
    f = lambda r, k: float((r.get(k) or "0").replace(",", "") or 0)
    tot_inst = sum(f(r, "Instructions Executed") for r in rows)
    tot_samp = sum(f(r, "# Samples") for r in rows)
    tot_wave = sum(f(r, "L1 Wavefronts Shared") for r in rows)
    tot_exc = sum(f(r, "L1 Wavefronts Shared Excessive") for r in rows)
    print(f"warp instructions {tot_inst:.0f}; stall samples {tot_samp:.0f}; shared wavefronts {tot_wave:.0f} (excessive {tot_exc:.0f})")
    by_op = defaultdict(lambda: [0.0, 0.0])
    for r in rows:
        op = r["Source"].split()[0] if r["Source"].split() else "?"
        if op.startswith("@"):
            op = r["Source"].split()[1]
        op = op.split(".")[0]
        by_op[op][0] += f(r, "Instructions Executed")
        by_op[op][1] += f(r, "# Samples")
    print("\nopcode            inst%   samples%")
    for op, (a, b) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:18]:
        print(f"{op:16s} {100 * a / tot_inst:6.2f}   {100 * b / max(tot_samp, 1):6.2f}")
    print("\ntop SASS by excessive shared wavefronts")
    for r in sorted(rows, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:10]:
        print(f"  {f(r, 'L1 Wavefronts Shared Excessive'):12.0f} / {f(r, 'L1 Wavefronts Shared'):12.0f}  {r['Source'][:70]}")
    print("\ntop SASS by stall samples")
    for r in sorted(rows, key=lambda r: -f(r, "# Samples"))[:top]:
        st = {k: f(r, k) for k in r if k.startswith("stall_") and "Not Issued" not in k and f(r, k) > 0}
        main_st = ", ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"  {f(r, '# Samples'):8.0f}  {r['Source'][:60]:60s} {main_st}")


if __name__ == "__main__":
    main()


def by_source_line(rep, top=40):
    """Instruction / stall-sample share per CUDA source line (needs -lineinfo + --import-source on)."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur_file, hdr, rows = None, None, []
    for l in out.splitlines():
        if l.startswith('"File Path"'):
            cur_file = next(csv.reader([l]))[1]
            hdr = None
            continue
        if l.startswith('"Line No"'):
            hdr = next(csv.reader([l]))
            continue
        if hdr and l.startswith('"') and not l.startswith('"Function Name"'):
            r = next(csv.reader([l]))
            if len(r) == len(hdr) and r[0]:          # source-line rows carry the line number; SASS rows have it empty
                d = {"file": cur_file, "line": r[0], "src": r[1]}
                for k, v in zip(hdr[2:], r[2:]):
                    d[k] = v
                rows.append(d)
    f = lambda d, k: float((d.get(k) or "0").replace(",", "") or 0)
    tot = sum(f(d, "Instructions Executed") for d in rows) or 1
    ts = sum(f(d, "# Samples") for d in rows) or 1
    print(f"\nper source line (total warp instructions {tot:.0f}, samples {ts:.0f})")
    for d in sorted(rows, key=lambda d: -f(d, "Instructions Executed"))[:top]:
        print(f'{d["file"].split("/")[-1][:20]:20s} L{d["line"]:>4} inst {100 * f(d, "Instructions Executed") / tot:5.2f}% '
              f'samp {100 * f(d, "# Samples") / ts:5.2f}%  {d["src"].strip()[:88]}')


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "lines":
    by_source_line(sys.argv[1], int(sys.argv[2]))
