"""Instruction mix of every loop of one kernel.  usage: sassloops.py <lib.so> <substring of the mangled kernel name>"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if sys.argv[2] not in name:
        continue
    ins = []
    for l in b.split("\n"):
        m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, len(ins), "instructions")
    for a, t in ins:
        m = re.search(r'BRA\S*\s+.*0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            c = collections.Counter(re.sub(r'@!?U?P\d+\s+', '', x[1]).split()[0].split('.')[0] for x in body)
            print("  loop %#x..%#x  %d instr: %s" % (tgt, a, len(body), dict(c.most_common())))
