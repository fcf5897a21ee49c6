"""Join the SASS source page of an ncu report with nvdisasm line info: executed instructions and stall samples per
CUDA source line / per opcode.

  cuobjdump -xelf all hcorepp_b200/lib/libhcore_b200.so && nvdisasm -g hcb_capi.sm_100a.cubin > all.sass
  ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
  python scripts/tools/ncu_source_join.py src.csv all.sass k_jacobi_svd_rxId [kernel index in the report]

The two listings hold the same instructions in the same order, so they are joined by position."""
import collections
import csv
import re
import sys


def disasm(path, func):
    lines = open(path).read().split("\n")
    st = next(i for i, l in enumerate(lines) if l.startswith(".text.") and func in l)
    en = next((i for i in range(st + 1, len(lines)) if lines[i].startswith("//-") and ".text." in lines[i]), len(lines))
    cur, ins = None, []
    for l in lines[st:en]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2).strip(), cur))
    return ins


def report(path, which):
    kern, c = [], None
    with open(path) as f:
        for row in csv.reader(f):
            if row and row[0] == "Kernel Name":
                c = {"name": row[1], "rows": []}
                kern.append(c)
            elif row and row[0] == "Address":
                c["hdr"] = row
            elif c is not None and len(row) > 10:
                c["rows"].append(row)
    return kern[which]


def main():
    src, sass, func = sys.argv[1:4]
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    ins, k = disasm(sass, func), report(src, which)
    ix = {h: i for i, h in enumerate(k["hdr"])}
    assert len(ins) == len(k["rows"]), (len(ins), len(k["rows"]))
    stalls = [h for h in k["hdr"] if h.startswith("stall_") and "Not Issued" not in h]
    by_op, by_line = collections.Counter(), collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
    op_s = collections.Counter()
    tot = ts = 0
    for (txt, loc), row in zip(ins, k["rows"]):
        n, s = int(row[ix["Instructions Executed"]]), int(row[ix["# Samples"]])
        m = re.match(r"(@!?U?P[T\d]+\s+)?([A-Z0-9_]+)", txt)
        op = m.group(2) if m else "?"
        by_op[op] += n
        op_s[op] += s
        a = by_line[loc]
        a[0] += n
        a[1] += s
        a[2][op] += n
        for h in stalls:
            v = int(row[ix[h]])
            if v:
                a[3][h[6:]] += v
        tot += n
        ts += s
    print(k["name"], "| executed warp instructions", tot, "| stall samples", ts)
    print("\n== by opcode: share of executed instructions / of stall samples")
    for op, n in by_op.most_common(24):
        print(f"{op:10s} {n / tot * 100:6.2f}%  {op_s[op] / ts * 100:6.2f}%")
    print("\n== by source line, sorted by executed instructions (top opcodes in % of all instructions)")
    for loc, a in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"{loc[0]}:{loc[1]:4d} instr {a[0] / tot * 100:5.2f}% samples {a[1] / ts * 100:5.2f}%  ",
              " ".join(f"{o}:{n / tot * 100:.1f}" for o, n in a[2].most_common(4)))
    print("\n== by source line, sorted by stall samples (dominant stall reasons in % of all samples)")
    for loc, a in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{loc[0]}:{loc[1]:4d} samples {a[1] / ts * 100:5.2f}%  ", " ".join(f"{o}:{n / ts * 100:.1f}" for o, n in a[3].most_common(4)))


if __name__ == "__main__":
    main()
