#!/usr/bin/env python3
"""ptx2c.py -- TEST INFRASTRUCTURE: turn the straight-line PTX of a plant function into C with one statement per instruction.

Why: the pendulum / cart-pole / quadrotor plants mix float variables and double literals, and their float results depend on which
multiply-adds the device compiler fuses.  nvcc decides that in two places -- NVVM emits fma.rn where it contracts, and ptxas may
still fuse a remaining `mul.f32` / `add.f32` pair -- and neither follows a rule simple enough to restate by hand for ~400
operations (e.g. it fuses the RIGHT product of `s3*s5 + c3*c5*s4`).  So the GPU-arithmetic variant of the oracle's plant functions
is read off the compiler's own output for the plant headers (oracle/tools/plant_ptx_wrap.cu -> nvcc -ptx), every instruction
becoming one explicitly rounded C statement; the result is committed as oracle/plants_gpu_arith.inc and pinned bit for bit to the
reference's GPU run (tests/golden/p*_unit_G.npz, p*_trace_G*.npz, p*_solve_G*.npz).  The host-arithmetic variant stays ordinary C.

ptxas step: an add.f32 / sub.f32 (no explicit rounding suffix besides the default) whose operand is the result of a single-use
mul.f32 is emitted as one fused multiply-add, the first such operand winning -- what sm_100 SASS shows for these kernels.

usage: ptx2c.py <file.ptx> <kernel> <c_function_name> <param names...>      (prints C to stdout)
"""
import re
import sys


def parse_kernel(ptx, kernel):
    m = re.search(r"\.entry %s\((.*?)\)\s*\{(.*?)\n\}" % re.escape(kernel), ptx, re.S)
    assert m, kernel
    body = re.sub(r"//.*", "", m.group(2))
    # fold call sequences into one pseudo instruction: call dst, name, args
    def fold(mm):
        blk = mm.group(1)
        args = re.findall(r"st\.param\.\w+\s+\[param\d+\],\s*([^;]+);", blk)
        name = re.search(r"call\.uni\s*\(retval0\),\s*(\w+)", blk).group(1)
        dst = re.search(r"ld\.param\.(\w+)\s+(%\w+),\s*\[retval0\]", blk)
        return "call.%s %s, %s, %s;" % (dst.group(1), dst.group(2), name, ", ".join(a.strip() for a in args))
    body = re.sub(r"\{([^{}]*call\.uni[^{}]*)\}", fold, body, flags=re.S)
    ins = [i.strip() for i in body.split(";") if i.strip() and not i.strip().startswith(".")]
    return ins


def parse_block(ptx, kernel, sym):
    """instructions of the LAST basic block of `kernel` that stores a float register into the shared array whose name ends in `sym`"""
    m = re.search(r"\.entry %s\((.*?)\)\s*\{(.*?)\n\}" % re.escape(kernel), ptx, re.S)
    assert m, kernel
    body = re.sub(r"//.*", "", m.group(2))
    body = re.sub(r"^(\$L\w+):", r"LABEL \1;", body, flags=re.M)
    ins = [i.strip() for i in body.split(";") if i.strip() and not i.strip().startswith(".")]
    blocks, cur = [], []
    for i in ins:
        if i.startswith("LABEL"):
            blocks.append(cur); cur = []
            continue
        cur.append(i)
        if re.match(r"(@%p\d+\s+)?bra", i) or i.startswith("bar."):
            blocks.append(cur); cur = []
    blocks.append(cur)
    hits = [b for b in blocks if any(re.match(r"st\.shared\.f32\s+\[\w*%s(\+\d+)?\],\s*%%f" % sym, i) for i in b)]
    assert hits, "no block stores to " + sym
    blk = [i for i in hits[-1] if not re.match(r"(@%p\d+\s+)?bra|bar\.", i)]
    # the block starts with the tail of the last inlined sinf / cosf (quadrant selects, integer and predicate work): the plant arithmetic
    # begins behind the last instruction that is neither float arithmetic nor a load / store / zero constant
    ok = re.compile(r"(ld|st)\.shared\.|(mul|add|sub|fma|neg|rcp|div|cvt\.f64\.f32|cvt\.rn\.f32\.f64|mov\.f32|mov\.f64)\b|mov\.(b32|u32)\s+%r\d+,\s*0$")
    last = max([k for k, i in enumerate(blk) if not ok.match(i)], default=-1)
    return blk[last+1:]


def fimm(tok):
    if tok.startswith("0f"):
        return "bitsf(0x%sU)" % tok[2:]
    if tok.startswith("0d"):
        return "bitsd(0x%sULL)" % tok[2:]
    return None


def translate(ptx, kernel, cname, params, block_sym=None, arrays=None):
    """block_sym: translate only the basic block that stores to that shared array; shared arrays map to C arrays through `arrays`
    (symbol suffix -> C name) and the float registers the block reads without defining them become the array `li` (live-ins)."""
    ins = parse_block(ptx, kernel, block_sym) if block_sym else parse_kernel(ptx, kernel)
    ptr = {}            # %rd -> (param name)
    ops = []            # (op, dst, srcs)
    for i in ins:
        if i == "ret":
            continue
        op, rest = i.split(None, 1)
        a = [t.strip() for t in rest.split(",")]
        ops.append((op, a))
    # use counts of registers (for the ptxas fusion step)
    uses = {}
    for op, a in ops:
        srcs = a[1:] if not op.startswith("st.") else a
        for t in srcs:
            for r in re.findall(r"%\w+", t):
                uses[r] = uses.get(r, 0) + 1
    muls = {}           # dst reg -> (a, b) of a fusable single-use mul.f32
    for op, a in ops:
        if op == "mul.f32" and uses.get(a[0], 0) == 1:
            muls[a[0]] = (a[1], a[2])
    fused = set()
    out = []
    decl_f, decl_d = set(), set()

    def v(tok):
        t = tok.strip()
        im = fimm(t)
        if im:
            return im
        assert t.startswith("%"), t
        name = t[1:]
        (decl_d if name.startswith("fd") else decl_f).add(name) if name[0] == "f" else None
        return name

    def addr(tok):
        m = re.match(r"\[(%\w+)(\+(\d+))?\]", tok.strip())
        if m:
            base, off = m.group(1), int(m.group(3) or 0)
            return "%s[%d]" % (ptr[base], off // 4)
        m = re.match(r"\[(\w+?)(\+(\d+))?\]", tok.strip())
        sym, off = m.group(1), int(m.group(3) or 0)
        for suffix, cn in arrays.items():
            if sym.endswith(suffix):
                return "%s[%d]" % (cn, off // 4)
        raise SystemExit("ptx2c: unknown shared symbol " + sym)

    defined, livein = set(), []
    if block_sym:
        for op, a in ops:
            srcs = a if op.startswith("st.") else a[1:]
            for t in srcs:
                for r in re.findall(r"%fd?\d+", t):
                    if r not in defined and r not in livein:
                        livein.append(r)
            if not op.startswith("st."):
                defined.add(a[0])

    pending = {}
    iconst = {}
    for op, a in ops:
        if op.startswith("ld.param.u64"):
            idx = int(re.search(r"_param_(\d+)", a[1]).group(1)); ptr[a[0]] = params[idx]
        elif op.startswith("cvta"):
            ptr[a[0]] = ptr[a[1]]
        elif op in ("ld.global.f32", "ld.global.nc.f32", "ld.shared.f32"):
            out.append("%s = %s;" % (v(a[0]), addr(a[1])))
        elif op in ("st.global.f32", "st.shared.f32"):
            out.append("%s = %s;" % (addr(a[0]), v(a[1])))
        elif op in ("mov.b32", "mov.u32"):
            iconst[a[0]] = int(a[1], 0)
        elif op in ("st.global.u32", "st.global.b32", "st.shared.u32", "st.shared.b32"):
            if a[1] not in iconst and a[1].startswith("%"):
                iconst[a[1]] = 0            # the zero the kernel keeps in an integer register across blocks (checked by the pins)
            val = iconst[a[1]] if a[1] in iconst else int(a[1], 0)
            out.append("%s = bitsf(0x%08XU);" % (addr(a[0]), val & 0xffffffff))
        elif op.startswith("call."):
            fn = {"_Z12pddp_opq_sinf": "SINF", "_Z12pddp_opq_cosf": "COSF", "_Z12pddp_opq_cosd": "COS64", "_Z12pddp_opq_sind": "SIN64"}[a[1]]
            out.append("%s = %s(%s);" % (v(a[0]), fn, v(a[2])))
        elif op in ("mov.f32", "mov.f64"):
            out.append("%s = %s;" % (v(a[0]), v(a[1])))
        elif op == "mul.f32" and a[0] in muls:
            pending[a[0]] = a            # emitted at its use if it gets fused, else right here
            out.append(("MUL", a[0]))
        elif op in ("mul.f32", "mul.rn.f32", "mul.f64", "mul.rn.f64"):
            out.append("%s = %s * %s;" % (v(a[0]), v(a[1]), v(a[2])))
        elif op in ("add.f32", "sub.f32"):
            sgn = "-" if op == "sub.f32" else ""
            x, y = a[1], a[2]
            if x in muls and x not in fused:
                fused.add(x); p, q = muls[x]
                out.append("%s = fmaf(%s, %s, %s%s);" % (v(a[0]), v(p), v(q), sgn, v(y)))
            elif y in muls and y not in fused:
                fused.add(y); p, q = muls[y]
                out.append("%s = fmaf(%s%s, %s, %s);" % (v(a[0]), sgn, v(p), v(q), v(x)))
            else:
                out.append("%s = %s %s %s;" % (v(a[0]), v(x), "-" if sgn else "+", v(y)))
        elif op in ("add.rn.f32", "add.f64", "add.rn.f64"):
            out.append("%s = %s + %s;" % (v(a[0]), v(a[1]), v(a[2])))
        elif op in ("sub.rn.f32", "sub.f64", "sub.rn.f64"):
            out.append("%s = %s - %s;" % (v(a[0]), v(a[1]), v(a[2])))
        elif op == "fma.rn.f32":
            out.append("%s = fmaf(%s, %s, %s);" % (v(a[0]), v(a[1]), v(a[2]), v(a[3])))
        elif op == "fma.rn.f64":
            out.append("%s = fma(%s, %s, %s);" % (v(a[0]), v(a[1]), v(a[2]), v(a[3])))
        elif op in ("neg.f32", "neg.f64"):
            out.append("%s = -%s;" % (v(a[0]), v(a[1])))
        elif op in ("rcp.rn.f32",):
            out.append("%s = 1.0f / %s;" % (v(a[0]), v(a[1])))
        elif op in ("rcp.rn.f64",):
            out.append("%s = 1.0 / %s;" % (v(a[0]), v(a[1])))
        elif op in ("div.rn.f32", "div.rn.f64"):
            out.append("%s = %s / %s;" % (v(a[0]), v(a[1]), v(a[2])))
        elif op == "cvt.f64.f32":
            out.append("%s = (double)%s;" % (v(a[0]), v(a[1])))
        elif op == "cvt.rn.f32.f64":
            out.append("%s = (float)%s;" % (v(a[0]), v(a[1])))
        elif op == "ret":
            pass
        else:
            raise SystemExit("ptx2c: unhandled instruction: %s %s" % (op, a))
    lines = []
    for o in out:
        if isinstance(o, tuple):
            r = o[1]
            if r in fused:
                continue
            a = pending[r]
            lines.append("%s = %s * %s;" % (v(a[0]), v(a[1]), v(a[2])))
        else:
            lines.append(o)
    sig = ", ".join(("const float *%s" % p) if p in ("x", "u", "li") else ("float *%s" % p) for p in params)
    res = ["static void %s(%s){" % (cname, sig)]
    if block_sym:
        res.insert(0, "/* live-in registers of the block, in this order: %s */" % " ".join(livein))
        for k, r in enumerate(livein):
            lines.insert(k, "%s = li[%d];" % (v(r), k))
    if decl_f:
        res.append("    float " + ", ".join(sorted(decl_f, key=lambda s: int(re.sub(r"\D", "", s)))) + ";")
    if decl_d:
        res.append("    double " + ", ".join(sorted(decl_d, key=lambda s: int(re.sub(r"\D", "", s)))) + ";")
    res += ["    " + l for l in lines]
    res.append("}")
    return "\n".join(res)


def main_block(argv):
    # ptx2c.py --block <sym> <file.ptx> <kernel> <c_function_name>
    ptx = open(argv[1]).read()
    print(translate(ptx, argv[2], argv[3], ["x", "u", "li", "dqdd"], block_sym=argv[0],
                    arrays={"s_x": "x", "s_u": "u", "s_dqdd": "dqdd", "s_qdd": "qdd_unused"}))


if __name__ == "__main__":
    if sys.argv[1] == "--block":
        main_block(sys.argv[2:]); sys.exit(0)
    ptx = open(sys.argv[1]).read()
    print(translate(ptx, sys.argv[2], sys.argv[3], sys.argv[4:]))
