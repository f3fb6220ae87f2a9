#!/usr/bin/env python3
"""Compare the SASS of two builds of libgpx.so function by function (cuobjdump -sass).

A refactor that is meant to change no instruction -- moving a helper between headers, sharing a body between two
kernels -- can be admitted without a GPU when every pre-existing function's SASS is byte-identical before and after
(clone suffixes like `$33` are renumbered when functions are added; bodies are compared with the suffix masked).

    python tools/sass_diff.py old/libgpx.so gigapaxos_b200/libgpx.so      # exit 1 if a common function differs
    python tools/sass_diff.py --hashes gigapaxos_b200/libgpx.so            # name -> sha1 of its SASS, as JSON
"""
import hashlib
import json
import re
import subprocess
import sys


def functions(lib: str) -> dict:
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, _, body = part.partition("\n")
        out[re.sub(r"\$\d+$", "$N", name.strip())] = re.sub(r"\$\d+", "$N", body)
    return out


def main(argv):
    if len(argv) == 3 and argv[1] == "--hashes":
        print(json.dumps({k: hashlib.sha1(v.encode()).hexdigest() for k, v in sorted(functions(argv[2]).items())}, indent=1))
        return 0
    if len(argv) != 3:
        print(__doc__)
        return 2
    a, b = functions(argv[1]), functions(argv[2])
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    print(f"{len(a)} functions before, {len(b)} after; added {sorted(set(b) - set(a))}; removed {sorted(set(a) - set(b))}")
    print("changed:", changed if changed else "none")
    return 1 if changed or set(a) - set(b) else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
