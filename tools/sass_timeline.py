"""Single-warp issue-model timeline of a straight-line SASS path (development tool, CPU only).

    cuobjdump -sass -fun <mangled> more4d_b200/csrc/attention.o > /tmp/k.sass
    python tools/sass_timeline.py /tmp/k.sass 0x2b70-0x2bc0 0x37a0-0x38b0 ...

Implements the in-order issue model of /opt/skills/guides/B300_MICROARCH.md ("Single-warp issue
model"): T = max(T + stall, scoreboards in wait_mask); variable-latency ops arm their write
scoreboard.  Control fields are decoded from the 128-bit encoding (stall [105:109), wbar [110:113),
rbar [113:116), wait_mask [116:122)).  The MUFU is modelled as a pipe that accepts one warp
instruction per 8 clocks with an 18-clock result latency; LDTM / STTM / SYNCS get nominal latencies.
Used to compare SCHEDULES of the softmax inner loop offline (no GPU in the authoring container);
absolute numbers are indicative only.
"""
import re
import sys

LAT = {"MUFU": 22, "LDTM": 60, "STTM": 20, "SYNCS": 30, "LDS": 29, "LDG": 400, "LDC": 30, "S2UR": 20, "R2UR": 12,
       "LDCU": 30, "VOTE": 10, "VOTEU": 10, "LDL": 40, "SHFL": 24}
MUFU_RT = 8


def parse(path):
    ins = []
    cur = None
    for line in open(path):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
        if m:
            cur = {"addr": int(m.group(1), 16), "text": m.group(2).strip(), "lo": int(m.group(3), 16)}
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", line)
        if m and cur is not None:
            enc = (int(m.group(1), 16) << 64) | cur["lo"]
            cur["stall"] = (enc >> 105) & 0xF
            cur["wbar"] = (enc >> 110) & 7
            cur["rbar"] = (enc >> 113) & 7
            cur["wait"] = (enc >> 116) & 0x3F
            ins.append(cur)
            cur = None
    return ins


def opclass(text):
    t = re.sub(r"^@!?U?P\d+\s+", "", text)
    return t.split()[0].split(".")[0]


def simulate(ins, verbose=False):
    T = 0
    sb = [0] * 6
    mufu_free = 0
    hist = {}
    t_first_mufu = t_last_mufu = None
    for i in ins:
        op = opclass(i["text"])
        arm = max([sb[s] for s in range(6) if i["wait"] >> s & 1], default=0)
        T = max(T, arm)
        if op == "MUFU":
            T = max(T, mufu_free)
            mufu_free = T + MUFU_RT
            t_first_mufu = T if t_first_mufu is None else t_first_mufu
            t_last_mufu = T
        if i["wbar"] < 6:
            sb[i["wbar"]] = max(sb[i["wbar"]], T + LAT.get(op, 20))
        if i["rbar"] < 6:
            sb[i["rbar"]] = max(sb[i["rbar"]], T + 6)
        if verbose:
            print(f"{T:6d}  {i['addr']:#06x} st={i['stall']:2d} w={i['wait']:02x} wb={i['wbar']} {i['text'][:70]}")
        hist[op] = hist.get(op, 0) + 1
        T += max(1, i["stall"])
    return T, hist, t_first_mufu, t_last_mufu


def main():
    ins = parse(sys.argv[1])
    by_addr = {i["addr"]: n for n, i in enumerate(ins)}
    verbose = "-v" in sys.argv
    path = []
    for seg in [a for a in sys.argv[2:] if not a.startswith("-")]:
        a, b = (int(x, 16) for x in seg.split("-"))
        path += ins[by_addr[a]:by_addr[b] + 1]
    T, hist, f, l = simulate(path, verbose)
    print(f"{len(path)} instructions, {T} clocks; first/last MUFU issue at {f}/{l}")
    print(" ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
