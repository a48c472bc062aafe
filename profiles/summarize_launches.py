"""Per-kernel mean duration of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/summarize_launches.py profiles/r1_launches_c2_v10.csv [iteration_kernel_name]"""
import collections
import csv
import re
import sys


def summarize(fn):
    rows = list(csv.reader(open(fn)))
    hdr, d = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            try:
                v = float(rec["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            name = re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "")
            unit = rec["Metric Unit"]
            v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
            d.setdefault(name, []).append(v)
    return d


if __name__ == "__main__":
    d = summarize(sys.argv[1])
    per = sys.argv[2] if len(sys.argv) > 2 else "iteration_update_kernel"
    for k, v in d.items():
        print(f"{k:44s} n={len(v):5d} mean={sum(v) / len(v):9.2f} us total={sum(v) / 1000:9.3f} ms")
    its = len(d.get(per, [])) or 1
    print(f"sum per {per}: {sum(sum(v) for v in d.values()) / its:.1f} us over {its} iterations")
