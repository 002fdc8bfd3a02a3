"""Function-level coverage of the package under the test-suite, without third-party tooling
(sys.monitoring, Python >= 3.12).  Off unless ``TDP_COV_DIR`` is set:

    TDP_COV_DIR=build/cov python -m pytest tests -q -m "not gpu"
    python tests/_cov.py build/cov            # functions of the package that never ran

Spawned ranks (tests/_mp.py) record into the same directory."""
import ast
import json
import os
import sys

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torchdistpackage_b200")
_seen = set()


def start() -> None:
    out = os.environ.get("TDP_COV_DIR")
    if not out or not hasattr(sys, "monitoring"):
        return
    mon = sys.monitoring
    try:
        mon.use_tool_id(mon.COVERAGE_ID, "tdp-func-cov")
    except ValueError:
        return

    def on_start(code, offset):
        if code.co_filename.startswith(PKG):
            _seen.add((os.path.relpath(code.co_filename, PKG), code.co_qualname))
        return mon.DISABLE

    mon.register_callback(mon.COVERAGE_ID, mon.events.PY_START, on_start)
    mon.set_events(mon.COVERAGE_ID, mon.events.PY_START)


def dump() -> None:
    out = os.environ.get("TDP_COV_DIR")
    if not out or not _seen:
        return
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"{os.getpid()}.json"), "w") as f:
        json.dump(sorted(_seen), f)


def _functions():
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            full = os.path.join(dirpath, fn)
            rel = os.path.relpath(full, PKG)
            tree = ast.parse(open(full).read())

            def walk(node, prefix):
                for ch in ast.iter_child_nodes(node):
                    if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef)):
                        q = prefix + ch.name
                        yield rel, q, ch.lineno
                        yield from walk(ch, q + ".<locals>.")
                    elif isinstance(ch, ast.ClassDef):
                        yield from walk(ch, prefix + ch.name + ".")
                    else:
                        yield from walk(ch, prefix)
            yield from walk(tree, "")


if __name__ == "__main__":
    d = sys.argv[1]
    seen = set()
    for f in os.listdir(d):
        if f.endswith(".json"):
            seen |= {tuple(x) for x in json.load(open(os.path.join(d, f)))}
    funcs = sorted(set(_functions()))
    missed = [(r, q, ln) for r, q, ln in funcs if (r, q) not in seen]
    print(f"{len(funcs) - len(missed)} / {len(funcs)} functions ran")
    for r, q, ln in missed:
        print(f"  {r}:{ln}  {q}")
