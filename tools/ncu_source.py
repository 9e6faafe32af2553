"""Per-SASS-instruction execution counts and stall samples from an .ncu-rep (needs -lineinfo / --import-source on)."""
import csv
import subprocess
import sys


def main(path, min_frac=0.002):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        try:
            data.append((int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), r[ix["Source"]].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data)
    tots = sum(d[1] for d in data)
    print(f"# total warp instructions {tot}, stall samples {tots}")
    for n, s, src in data:
        if n >= min_frac * tot or s >= min_frac * tots:
            print(f"{n:12d} {100 * n / tot:5.1f}% {100 * s / max(tots, 1):5.1f}%s  {src}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.002)
