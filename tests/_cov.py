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
    events = mon.events.PY_START
    if os.environ.get("TDP_COV_LINES"):                 # line mode: also which lines ran
        def on_line(code, line):
            if code.co_filename.startswith(PKG):
                _seen.add((os.path.relpath(code.co_filename, PKG), int(line)))
            return mon.DISABLE
        mon.register_callback(mon.COVERAGE_ID, mon.events.LINE, on_line)
        events |= mon.events.LINE
    mon.set_events(mon.COVERAGE_ID, events)


def dump() -> None:
    out = os.environ.get("TDP_COV_DIR")
    if not out or not _seen:
        return
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"{os.getpid()}.json"), "w") as f:
        json.dump(sorted(_seen, key=lambda t: (t[0], str(t[1]))), f)


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


def _executable_lines(path):
    """Line numbers that carry code, from the compiled code objects (docstrings excluded)."""
    lines = set()
    todo = [compile(open(path).read(), path, "exec")]
    while todo:
        co = todo.pop()
        doc = co.co_consts[0] if co.co_consts and isinstance(co.co_consts[0], str) else None
        for _, _, ln in co.co_lines():
            if ln is not None:
                lines.add(ln)
        todo += [c for c in co.co_consts if hasattr(c, "co_lines")]
    return lines


def _line_report(seen, only):
    ran = {}
    for item in seen:
        if isinstance(item[1], int):
            ran.setdefault(item[0], set()).add(item[1])
    for rel in sorted(ran):
        if only and not any(o in rel for o in only):
            continue
        src = open(os.path.join(PKG, rel)).read().splitlines()
        todo = sorted(_executable_lines(os.path.join(PKG, rel)) - ran[rel])
        todo = [ln for ln in todo if not src[ln - 1].lstrip().startswith(('"""', "def ", "class ", "@"))]
        print(f"== {rel}: {len(todo)} executable lines never ran")
        start = prev = None
        for ln in todo + [None]:
            if start is None:
                start = prev = ln
            elif ln is not None and ln <= prev + 2:
                prev = ln
            else:
                print(f"   {start}-{prev}: {src[start - 1].strip()[:90]}")
                start = prev = ln


if __name__ == "__main__":
    d = sys.argv[1]
    seen = set()
    for f in os.listdir(d):
        if f.endswith(".json"):
            seen |= {tuple(x) for x in json.load(open(os.path.join(d, f)))}
    if len(sys.argv) > 2 and sys.argv[2] == "--lines":
        _line_report(seen, sys.argv[3:])
        sys.exit(0)
    seen = {x for x in seen if isinstance(x[1], str)}
    funcs = sorted(set(_functions()))
    missed = [(r, q, ln) for r, q, ln in funcs if (r, q) not in seen]
    print(f"{len(funcs) - len(missed)} / {len(funcs)} functions ran")
    for r, q, ln in missed:
        print(f"  {r}:{ln}  {q}")
