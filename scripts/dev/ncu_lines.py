"""Per-source-line summary of an ncu report (needs --import-source on and -lineinfo): executed warp-instructions and
warp-stall samples by CUDA source line.  usage: python scripts/dev/ncu_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                         text=True).stdout
    cur, key = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    src = {}
    i_s = i_e = None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print(r[1])
        elif r[0] == "Line No":
            i_s, i_e = r.index("# Samples"), r.index("Instructions Executed")
        elif r[0].strip():
            key = (cur, int(r[0]))
            src[key] = r[1].strip()[:110]
        elif key is not None:
            try:
                agg[key][0] += int(r[i_s] or 0)
                agg[key][1] += int(r[i_e] or 0)
                agg[key][2] += 1
            except (ValueError, IndexError):
                pass
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_e = sum(v[1] for v in agg.values()) or 1
    print("samples", tot_s, "warp-instructions", tot_e)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] * 100 / tot_e:5.2f}% exec {v[0] * 100 / tot_s:5.2f}% samples sass={v[2]:4d}  {k[0]}:{k[1]}  {src.get(k, '')}")


if __name__ == "__main__":
    main()
