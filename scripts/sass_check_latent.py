"""Static proof that the packed-f32x2 latent kernels keep products and sums separately rounded.

ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false, so csrc/latent_core.cuh forms every
packed product as fma(a, b, -0.0) with the -0.0 read from constant memory (lat_negzero).  This script disassembles the built
object and checks, for every kernel whose name contains one of the given substrings, that the addend of EVERY FFMA2 is either a
uniform register / constant operand or a register pair whose only writers are loads from the constant bank (i.e. the opaque
-0.0), which makes each FFMA2 an exactly rounded product.  Writes a digest (JSON) for profiles/ and exits 1 on a violation.

usage: python scripts/sass_check_latent.py [object] [--out profiles/xxx.json] [--kernels r2]
"""
import argparse
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(obj, needles):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    report, ok = {}, True
    for name, ins in funcs.items():
        if not any(n in name for n in needles):
            continue
        ffma2 = [i for i in ins if re.search(r"\bFFMA2\b", i)]
        addend_regs, uniform = collections.Counter(), 0
        for i in ffma2:
            ops = [o.strip() for o in i.split("FFMA2", 1)[1].split(",")]
            c = ops[3]
            if c.startswith("UR") or c.startswith("c["):
                uniform += 1
            else:
                addend_regs[re.match(r"-?\|?(R\d+)", c).group(1)] += 1
        # nearest preceding writer (linear order) of each register addend must be a load of the constant -0.0
        violations, unresolved = [], []

        def writes(i, n):
            body = re.sub(r"^@!?U?P\d+\s+", "", i)
            parts = body.split(None, 1)
            if len(parts) < 2:
                return False
            op, dst = parts[0], parts[1].split(",")[0].strip().split(".")[0]
            if not re.fullmatch(r"R\d+", dst):
                return False
            d = int(dst[1:])
            wide = 2 if (".64" in op or op in ("FFMA2", "FADD2", "FMUL2", "DMUL", "DADD", "DFMA", "F2F.F64.F32")
                         or op.startswith("IMAD.WIDE")) else (4 if ".128" in op else 1)
            return d <= n + 1 and d + wide - 1 >= n

        for k, i in enumerate(ins):
            if not re.search(r"\bFFMA2\b", i):
                continue
            c = [o.strip() for o in i.split("FFMA2", 1)[1].split(",")][3]
            if c.startswith("UR") or c.startswith("c["):
                continue
            n = int(re.match(r"-?\|?R(\d+)", c).group(1))
            j, crossed = k - 1, False
            while j >= 0 and not writes(ins[j], n):
                crossed |= bool(re.match(r"(BRA|EXIT|RET|JMP)\b", ins[j]))  # block not entered by fall-through
                j -= 1
            src = ins[j] if j >= 0 else "<none>"
            if crossed:  # linear order says nothing here: the defining block is a branch target away
                unresolved.append({"ffma2": i, "linear_predecessor_block_writer": src})
                continue
            if not (re.search(r"\bLDC(\.64)?\b", src) and "c[0x3]" in src) and not re.search(r"\bMOV\b.*\bUR\d+", src) \
                    and not re.search(r"IMAD\.U32 R\d+, RZ, RZ, UR\d+", src):
                violations.append({"ffma2": i, "addend_written_by": src})
        report[name] = {"FFMA2": len(ffma2), "addend_uniform_or_constant": uniform,
                        "addend_register_pairs": dict(addend_regs), "non_constant_writers_of_addend": violations,
                        "not_decidable_in_linear_order": unresolved,
                        "FMUL2": sum(bool(re.search(r"\bFMUL2\b", i)) for i in ins),
                        "FADD2": sum(bool(re.search(r"\bFADD2\b", i)) for i in ins)}
        ok &= not violations and len(ffma2) > 0
    return ok, report


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("obj", nargs="?", default=os.path.join(ROOT, "waves.jl_b200", "csrc", "latent_abi.o"))
    ap.add_argument("--out")
    ap.add_argument("--kernels", nargs="*", default=["_r2"])
    a = ap.parse_args()
    ok, rep = check(a.obj, a.kernels)
    out = {"object": os.path.relpath(a.obj, ROOT), "every_FFMA2_addend_is_the_opaque_constant": ok, "kernels": rep}
    print(json.dumps(out, indent=1))
    if a.out:
        json.dump(out, open(a.out, "w"), indent=1)
    sys.exit(0 if ok else 1)
