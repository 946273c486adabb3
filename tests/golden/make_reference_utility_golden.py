"""Generates tests/golden/reference_get_nzd.json by EXECUTING the reference's own second formulation of fftFIT for the z
direction, `get_nzd` in /root/reference/utilities/in_helper.py:11-23 (the function is extracted from the file's text with
`ast`; the module itself cannot be imported: it needs the authors' external `channel` package and reads stdin).
Run in the build container, where /root/reference exists:  python tests/golden/make_reference_utility_golden.py"""
import ast
import copy
import json
import os

REF = "/root/reference/utilities/in_helper.py"


def reference_get_nzd(path=REF):
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_nzd")
    ns = {"copy": copy.copy}
    exec(compile(ast.Module([fn], []), path, "exec"), ns)
    return ns["get_nzd"]


if __name__ == "__main__":
    f = reference_get_nzd()
    nz = list(range(1, 2101))
    out = {"source": "get_nzd, utilities/in_helper.py:11-23 of davecats/channel, executed", "nz": nz, "nzd": [int(f(n)) for n in nz]}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_get_nzd.json")
    json.dump(out, open(dst, "w"))
    print(dst, out["nzd"][:12], out["nzd"][188], out["nzd"][1022])
