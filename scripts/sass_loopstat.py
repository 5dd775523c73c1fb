"""Static instruction counts of the Kalman time loops, read from the SASS of the built objects.

    python scripts/sass_loopstat.py            # -> profiles/r02_sass_loop_counts.json + loop excerpts (P = 5)

For every AR order P the two innermost loops of loglik_batch_kernel<P> (K1) are located (backward branch, >= 8
DFMA): the all-conjugate-pairs loop and the generic loop.  Per Kalman step: instructions of every kind, FP64
instructions by opcode, DFMAs that read three distinct register pairs without an operand-reuse flag (those hold
the dispatch port 3 cycles instead of 2 on sm_100a: profiles/r02a_fp64_issue_probe.txt), and the resulting
dispatch-cycle model  sum(FP64: 2 or 3) + (every other instruction: 1).  bench.py reads the JSON to report the
EXECUTED FP64 rate next to the algorithmic one of SURVEY section 8d.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "carma_pack_b200", "build")
OUT = os.path.join(ROOT, "profiles")


def disasm(obj, fun):
    return subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout


def parse(txt):
    ins = []
    for l in txt.splitlines():
        m = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def loops(ins, min_dfma=8):
    addr = {a: i for i, (a, _) in enumerate(ins)}
    found = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                if sum("DFMA" in x for _, x in body) >= min_dfma:
                    found.append(body)
    # innermost = loops that contain no other found loop
    inner = [b for b in found if not any(o is not b and o[0][0] >= b[0][0] and o[-1][0] <= b[-1][0] for o in found)]
    return inner


def stats(body):
    ops = collections.Counter()
    three = cycles = 0
    for _, t in body:
        tt = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = tt.split()[0].split(".")[0]
        ops[op] += 1
        if op in ("DFMA", "DMUL", "DADD", "DSETP"):
            args = tt[len(tt.split()[0]):].split(",")
            srcs = args[1:] if op != "DSETP" else args[2:]
            regs = set()
            for s_ in srcs:
                m = re.match(r"^[-|]*R(\d+)(\.reuse)?", s_.strip())
                if m and not m.group(2):
                    regs.add(m.group(1))
            three += len(regs) == 3
            cycles += max(2, len(regs))
        else:
            cycles += 1
    fp64 = {k: ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP")}
    return {"instructions": len(body), "fp64_instructions": sum(fp64.values()), "fp64": fp64,
            "fp64_flops": 2 * fp64["DFMA"] + fp64["DMUL"] + fp64["DADD"], "dfma_three_register_operands": three,
            "dispatch_cycle_model": cycles, "other": {k: v for k, v in ops.most_common() if k not in fp64}}


def main():
    os.makedirs(OUT, exist_ok=True)
    obj = os.path.join(BUILD, "loglik.cu.o")
    out = {"source": "cuobjdump -sass of carma_pack_b200/build/loglik.cu.o (nvcc 12.9, sm_100a)", "k1": {}}
    for P in range(1, 8):
        fun = "_ZN5carma19loglik_batch_kernelILi%dEEEvNS_10SeriesViewEiiij11carma_priorPKdPdm" % P
        ins = parse(disasm(obj, fun))
        inner = sorted(loops(ins), key=lambda b: len(b))
        hot = [b for b in inner if sum(("LDS" in t) for _, t in b) >= 3]   # the staged-series loops
        if not hot:
            continue
        entry = {"all_conjugate_loop": stats(hot[0])}
        if len(hot) > 1:
            entry["generic_loop"] = stats(hot[1])
        entry["kernel_instructions_total"] = len(ins)
        out["k1"][str(P)] = entry
        if P == 5:
            for name, b in (("all_conjugate", hot[0]),) + ((("generic", hot[1]),) if len(hot) > 1 else ()):
                with open(os.path.join(OUT, "r02_k1_loop_P5_%s.sass" % name), "w") as f:
                    f.write("// loglik_batch_kernel<5>: %s time loop, one Kalman step (cuobjdump -sass)\n" % name)
                    for a, t in b:
                        f.write("/*%05x*/  %s ;\n" % (a, t))
    # K3: the filter loops of logdensity_resident<5> (noinline, emitted inside pt_kernel<5, false>): all-conjugate and
    # generic, each plain and software-pipelined (the pipelined pair runs on launches of <= 296 blocks)
    obj3 = os.path.join(BUILD, "mcmc.cu.o")
    fun3 = "_ZN5carma9pt_kernelILi5ELb0EEEvNS_10SeriesViewENS_8PTParamsEmNS_7PTMultiE"
    ins3 = parse(disasm(obj3, fun3))
    hot3 = sorted([b for b in loops(ins3, min_dfma=40) if sum(("LDS" in t) for _, t in b) >= 3], key=len)
    out["k3"] = {"function": "pt_kernel<5, false> (logdensity_resident<5> inlined text)", "kernel_instructions_total": len(ins3),
                 "filter_loops_by_size": [stats(b) for b in hot3]}
    if hot3:
        # the generic plain loop is what config 3 executes (a warp mixes chains with and without real root pairs)
        base = stats(hot3[0])["fp64_instructions"]
        pick = next((b for b in hot3 if stats(b)["fp64_instructions"] > base + 5), hot3[0])
        with open(os.path.join(OUT, "r02_k3_loop_P5_generic.sass"), "w") as f:
            f.write("// pt_kernel<5,false>: generic time loop of logdensity_resident<5>, one Kalman step (cuobjdump -sass)\n")
            for a, t in pick:
                f.write("/*%05x*/  %s ;\n" % (a, t))
    with open(os.path.join(OUT, "r02_sass_loop_counts.json"), "w") as f:
        json.dump(out, f, indent=1)
    for P, e in out["k1"].items():
        a = e["all_conjugate_loop"]
        g = e.get("generic_loop", {})
        print("P=%s  all-conjugate: %d instr, %d FP64 (%d flops), model %d cycles | generic: %s instr, %s FP64" % (
            P, a["instructions"], a["fp64_instructions"], a["fp64_flops"], a["dispatch_cycle_model"],
            g.get("instructions"), g.get("fp64_instructions")))
    print("K3 P=5 loops (instr, FP64, model cycles):", [(x["instructions"], x["fp64_instructions"], x["dispatch_cycle_model"])
                                                        for x in out["k3"]["filter_loops_by_size"]])


if __name__ == "__main__":
    sys.exit(main())
