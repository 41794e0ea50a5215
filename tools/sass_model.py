"""Offline issue/operand-bandwidth model of the hot loop of a kernel, from its SASS.

    python tools/sass_model.py <file.o|.cubin|.so> <function-regex> [--dump]

Finds the innermost backward-branch loop with the most FMA work and reports, for that
loop body, the instruction mix and an estimate of the cycles one warp needs on its SM
sub-partition, using the measured B200 numbers (tools/microbench.cu, tools/pipes.cu):
  FFMA  1.05 cyc, FFMA2 2.0 cyc pipe time; MUFU 8.1 cyc of XU; 1 issue slot / instruction;
  register file: 2 banks (even / odd register index), one 32-bit read per bank per cycle;
  an operand marked .reuse on the previous instruction (same slot) costs no read.
"""
import re
import subprocess
import sys

INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);")


def disasm(path, fn_regex):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur, name = {}, None, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    sel = [k for k in funcs if re.search(fn_regex, k)]
    if not sel:
        raise SystemExit(f"no function matches {fn_regex}; have {list(funcs)[:5]}...")
    return sel[0], funcs[sel[0]]


def parse(text):
    pred = ""
    if text.startswith("@"):
        pred, text = text.split(None, 1)
    parts = text.split(None, 1)
    op = parts[0]
    args = [a.strip() for a in parts[1].split(",")] if len(parts) > 1 else []
    return pred, op, args


def reg_of(arg):
    m = re.match(r"[-|~!]*\|?R(\d+)", arg)
    return int(m.group(1)) if m else None


def model(body, dump=False):
    mix, issue, fma_pipe, xu, rf_stall = {}, 0, 0.0, 0.0, 0.0
    prev_reuse = {}
    fresh_hist = {}
    for addr, text in body:
        pred, op, args = parse(text)
        base = op.split(".")[0]
        mix[base] = mix.get(base, 0) + 1
        issue += 1
        srcs = args[1:] if args else []
        cost_pipe = 0.0
        if base in ("FFMA", "FMUL", "FADD"):
            cost_pipe = 1.05
        elif base in ("FFMA2", "FMUL2", "FADD2"):
            cost_pipe = 2.0
        if base == "MUFU":
            xu += 8.1
        # the reuse cache is per operand slot: an instruction that does not use a slot
        # leaves that slot's cached register alone (ptxas relies on this across MUFU/LDS)
        this_reuse = dict(prev_reuse)
        for slot, a in enumerate(srcs):
            if reg_of(a) is not None and ".reuse" not in a:
                this_reuse.pop(slot, None)
        if cost_pipe:
            wide = base.endswith("2")
            even = odd = 0
            seen = set()
            for slot, a in enumerate(srcs):
                r = reg_of(a)
                if r is None or "RZ" in a:
                    continue
                if ".reuse" in a:
                    this_reuse[slot] = r
                if prev_reuse.get(slot) == r or r in seen:
                    continue
                seen.add(r)
                if wide and ".F32x2" in a:      # 64-bit pair: one register in each bank
                    even += 1
                    odd += 1
                elif r % 2 == 0:              # scalar operand (also FFMA2's broadcast .F32 form)
                    even += 1
                else:
                    odd += 1
            reads = max(even, odd)
            key = (base, reads)
            fresh_hist[key] = fresh_hist.get(key, 0) + 1
            c = max(cost_pipe, float(reads))
            rf_stall += c - cost_pipe
            fma_pipe += c
            if dump:
                print(f"  {addr:05x} {text:70s} reads/bank={reads} cost={c:.2f}")
        elif dump:
            print(f"  {addr:05x} {text}")
        prev_reuse = this_reuse
    return mix, issue, fma_pipe, xu, rf_stall, fresh_hist


def find_hot_loop(ins):
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:\S+,\s*)?`?\(?(0x[0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addr_index:
            continue
        body = ins[addr_index[tgt]:i + 1]
        work = sum(2 if "FFMA2" in x else 1 for _, x in body if "FFMA" in x)
        # innermost: prefer the smallest body among those with near-maximal density
        score = work / (len(body) ** 0.5)
        if work >= 16 and (best is None or score > best[0]):
            best = (score, body)
    if best is None:
        raise SystemExit("no FMA loop found")
    return best[1]


def main():
    path, fn = sys.argv[1], sys.argv[2]
    dump = "--dump" in sys.argv
    name, ins = disasm(path, fn)
    body = find_hot_loop(ins)
    mix, issue, fma_pipe, xu, rf_stall, hist = model(body, dump)
    fmas = sum(v * (2 if k.endswith("2") else 1) for k, v in mix.items() if k.startswith("FFMA"))
    items = fmas / 17.0
    print(f"{name[:100]}")
    print(f"hot loop {body[0][0]:#x}..{body[-1][0]:#x}: {len(body)} instructions, FMAs {fmas} (~{items:.2f} items)")
    print("mix:", ", ".join(f"{k}={v}" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])))
    print("register reads per bank (op, reads): count ->", dict(sorted(hist.items())))
    # Measured on B200 (tools/pipes.cu): the dispatch port is what everything shares - an FFMA2
    # holds it for 2 cycles (3 when it needs 3 register reads from one bank), every other
    # instruction for 1 - so the serial sum below is the estimate, not the max over pipes.
    n_fma_instr = sum(v for k, v in mix.items() if k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
    serial = fma_pipe + (issue - n_fma_instr)
    print(f"per item: instructions {issue / items:.2f}  FMA-pipe {fma_pipe / items:.2f} (operand stalls "
          f"{rf_stall / items:.2f})  XU {xu / items:.2f}  -> dispatch-serial estimate {serial / items:.2f} cyc/item "
          f"= {100 * 18 * items / serial:.1f}% of FP32 peak")


if __name__ == "__main__":
    main()
