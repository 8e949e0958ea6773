"""Crude undefined-name check (no pyflakes in the image): names loaded somewhere in a module that nothing in the module binds.
python scripts/check_names.py file.py ..."""
import ast
import builtins
import sys


def check(path):
    tree = ast.parse(open(path).read())
    defined = set(dir(builtins)) | {"__file__"}
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
            defined.add(n.name)
        if isinstance(n, (ast.FunctionDef, ast.Lambda)):
            a = n.args
            for x in a.args + a.kwonlyargs + a.posonlyargs:
                defined.add(x.arg)
            if a.vararg:
                defined.add(a.vararg.arg)
            if a.kwarg:
                defined.add(a.kwarg.arg)
        if isinstance(n, (ast.Import, ast.ImportFrom)):
            for x in n.names:
                defined.add((x.asname or x.name).split(".")[0])
        if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            defined.add(n.id)
        if isinstance(n, ast.ExceptHandler) and n.name:
            defined.add(n.name)
    bad = 0
    for n in ast.walk(tree):
        if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in defined:
            print(f"{path}:{n.lineno}: undefined name {n.id}")
            bad += 1
    return bad


if __name__ == "__main__":
    sys.exit(1 if sum(check(p) for p in sys.argv[1:]) else 0)
