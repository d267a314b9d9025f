"""Condenses an `ncu --csv --metrics ...` launch list into one line per launch (kernel, metrics)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    launches = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        name = re.sub(r"<unnamed>::", "", name)[:60]
        launches.setdefault(r[0], {"kernel": name, "grid": r[8]})[r[-3]] = r[-1]
    for i, d in launches.items():
        k = d.pop("kernel")
        g = d.pop("grid")
        print(i, k, g, " ".join(f"{m.split('.')[0].replace('__', ':')}={v}" for m, v in d.items()))


if __name__ == "__main__":
    main(sys.argv[1])
