"""TEST / BASELINE INFRASTRUCTURE ONLY.  Parameter classes of every __global__ kernel of the reference's kernel file,
parsed from the source where it lies (/root/reference, never copied), written to oracle/_ref/ref_signatures.json: the
GPU stand-in for pycuda (oracle/ref_harness_gpu) marshals launch arguments positionally from it on machines that do not
have the reference source (the GPU box).   python oracle/make_signatures.py <kernel_sparse_adapt.cu> <out.json>"""
import json
import re
import sys


def signatures(path):
    src = re.sub(r"//[^\n]*", "", open(path).read())
    sigs = {}
    for m in re.finditer(r"__global__\s+void\s+(\w+)\s*\(([^)]*)\)", src):
        kinds = []
        for prm in m.group(2).split(","):
            prm = prm.strip()
            if not prm:
                continue
            if "*" in prm:
                kinds.append("p")
            elif re.search(r"\bfloat2\b", prm):
                kinds.append("x")
            elif re.search(r"\bfloat\b", prm):
                kinds.append("f")
            elif re.search(r"\bdouble\b", prm):
                kinds.append("d")
            elif "long" in prm:
                kinds.append("q")
            else:
                kinds.append("i")
        sigs[m.group(1)] = kinds
    return sigs


if __name__ == "__main__":
    json.dump(signatures(sys.argv[1]), open(sys.argv[2], "w"), indent=0, sort_keys=True)
    print("built", sys.argv[2])
