"""Developer tool (no GPU): decode the scheduling control bits of a kernel's SASS (write / read barrier = scoreboard slot an
instruction sets, wait mask = the slots it waits for) to find FALSE scoreboard dependences - an instruction in a loop waiting on
the slot of a load issued outside it, which loads issued inside the loop share.

    python tools/sass_waits.py OBJECT MANGLED_KERNEL_NAME [--slot N]

Prints every global load (with the slot it sets) and every instruction that waits on slot N (default: the slot of the loop's
LDG.U16 prefetch loads if there is one, else 5).  Control field (sm_7x .. sm_100): bits 105-108 stall count, 109 yield, 110-112
write barrier, 113-115 read barrier, 116-121 wait mask of the 128-bit instruction.

Round 2: the column-walk forward's first use of A (an FMUL2 in the window loop) waited on slot 5 = the slot of A's per-item LDG AND
of the B / C prefetch LDGs issued at the top of the same window: every window stalled for its own prefetch (ncu: 22 % of the warp
samples of the L = 6400 forward on that instruction).  profiles/r2zz_fwd_prefetch_wait.md.
"""
import re
import subprocess
import sys


def load(obj, fun):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True, check=True).stdout
    lines, ins, i = txt.splitlines(), [], 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]+) \*/", lines[i + 1])
            if m2:
                ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
                i += 2
                continue
        i += 1
    return ins


def ctrl(hi):
    c = hi >> 41
    return {"stall": c & 0xF, "yield": (c >> 4) & 1, "wbar": (c >> 5) & 7, "rbar": (c >> 8) & 7, "wait": (c >> 11) & 0x3F}


def main():
    obj, fun = sys.argv[1], sys.argv[2]
    ins = load(obj, fun)
    slot = None
    if "--slot" in sys.argv:
        slot = int(sys.argv[sys.argv.index("--slot") + 1])
    else:
        for _, t, hi in ins:
            if t.startswith("LDG.E.U16") and ctrl(hi)["wbar"] != 7:
                slot = ctrl(hi)["wbar"]
        slot = 5 if slot is None else slot
    print(f"# {fun}\n# slot {slot}: loads that set it, instructions that wait for it")
    for a, t, hi in ins:
        c = ctrl(hi)
        if "LDG" in t and "UTMALDG" not in t:
            print(f"{a:#07x}  {t[:84]:84s} sets slot {c['wbar'] if c['wbar'] != 7 else '-'}  waits {c['wait']:06b}")
        elif c["wait"] & (1 << slot):
            print(f"{a:#07x}  {t[:84]:84s}                waits {c['wait']:06b}")


if __name__ == "__main__":
    main()
