"""SASS / ptxas evidence of the hot kernels, generated on the CPU from the built library.

    python profiles/sass_summary.py          # writes profiles/sass/*.txt

For every hot kernel: the `ptxas -v` resource line (from csrc/*.ptxas.log), an opcode histogram of its SASS
(cuobjdump -sass of libjfem_b200.so, sm_100a) and the lines that prove the asynchronous / bulk / 128-bit paths
(LDGSTS.*.128, UBLKCP, SYNCS, USETMAXREG, LDG.E.128 ...).
"""
import collections
import glob
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "juliafem.jl_b200", "libjfem_b200.so")
HOT = [("patch_kernel_ws<10, 1, 0, PtLinear>", r"patch_kernel_wsILi10ELi1ELi0EN2jf8PtLinear"),
       ("patch_kernel_ws<8, 1, 0, PtLinear>", r"patch_kernel_wsILi8ELi1ELi0EN2jf8PtLinear"),
       ("patch_kernel<10, 0, 0, PtLinear, 256>", r"patch_kernelILi10ELi0ELi0EN2jf8PtLinearELi256"),
       ("iface_reduce_kernel", r"iface_reduce_kernel"),
       ("elem_warp_kernel<10, PtLinear>", r"elem_warp_kernelILi10EN2jf8PtLinear"),
       ("elem_warp_kernel<10, PtPPTangent>", r"elem_warp_kernelILi10EN2jf11PtPPTangent"),
       ("spmv_kernel", r"spmv_kernel"),
       ("cg_update_kernel", r"cg_update_kernel"),
       ("cg_dot_kernel", r"cg_dot_kernel"),
       ("cg_p_kernel", r"cg_p_kernel")]
PROOF = re.compile(r"LDGSTS|UBLKCP|SYNCS|USETMAXREG|LDG\.E\.128|LDS\.128|STS\.128|STG\.E\.128|UTMALDG|ACQBULK|MEMBAR|ATOM|RED\.")


def ptxas_lines():
    out = {}
    for f in glob.glob(os.path.join(ROOT, "juliafem.jl_b200", "csrc", "*.ptxas.log")):
        lines = open(f).read().splitlines()
        for i, ln in enumerate(lines):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if m:
                res = [x.strip() for x in lines[i + 1:i + 4] if "ptxas info" in x or "bytes stack" in x]
                out[m.group(1)] = " | ".join(res)
    return out


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    px = ptxas_lines()
    os.makedirs(os.path.join(HERE, "sass"), exist_ok=True)
    index = []
    for label, pat in HOT:
        hit = [f for f in funcs if re.search(pat, f.split("\n", 1)[0])]
        if not hit:
            print("not found:", label)
            continue
        body = hit[0]
        name = body.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
        hist = collections.Counter(i for i in ins)
        fam = collections.Counter(i.split(".")[0] for i in ins)
        proof = [ln.strip() for ln in body.splitlines() if PROOF.search(ln)]
        pcount = collections.Counter(re.search(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", re.sub(r"/\*[0-9a-f]+\*/", "", ln)).group(1) for ln in proof)
        fn = os.path.join(HERE, "sass", re.sub(r"[^A-Za-z0-9]+", "_", label).strip("_") + ".txt")
        with open(fn, "w") as fh:
            fh.write(f"{label}\nmangled: {name}\nptxas -v: {px.get(name, '(not in log)')}\n")
            fh.write(f"instructions: {len(ins)}\n\nopcode families (top 30):\n")
            for k, v in fam.most_common(30):
                fh.write(f"  {k:14s} {v}\n")
            fh.write("\nasynchronous / bulk / wide / synchronisation opcodes (full mnemonic: count):\n")
            for k, v in sorted(pcount.items()):
                fh.write(f"  {k:34s} {v}\n")
            fh.write("\nfp64: " + ", ".join(f"{k} {v}" for k, v in sorted(hist.items()) if k.startswith(("DFMA", "DADD", "DMUL", "DSETP", "MUFU.RCP64"))) + "\n")
            fh.write("\nfirst occurrences:\n")
            seen = set()
            for ln in proof:
                op = re.sub(r"/\*[0-9a-f]+\*/", "", ln).split()[0:2]
                key = " ".join(op)[:24]
                if key not in seen and len(seen) < 24:
                    seen.add(key)
                    fh.write("  " + ln[:150] + "\n")
        index.append((label, len(ins), px.get(name, "")))
        print(f"{label}: {len(ins)} instructions; {px.get(name, '')}")
    with open(os.path.join(HERE, "sass", "INDEX.txt"), "w") as fh:
        for label, n, p in index:
            fh.write(f"{label}: {n} SASS instructions; {p}\n")


if __name__ == "__main__":
    main()
