"""`compat/` on PYTHONPATH: a flow written against `bytewax` loads unchanged and runs on this engine."""

import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flow_written_against_bytewax_runs_unchanged(tmp_path):
    (tmp_path / "wc_flow.py").write_text(textwrap.dedent("""
        import bytewax.operators as op
        from bytewax.connectors.stdio import StdOutSink
        from bytewax.dataflow import Dataflow
        from bytewax.testing import TestingSource

        flow = Dataflow("wc")
        lines = op.input("inp", flow, TestingSource(["to be or not to be", "that is"]))
        words = op.flat_map("split", lines, str.split)
        op.output("out", op.count_final("count", words, lambda w: w), StdOutSink())
    """))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    r = subprocess.run([sys.executable, "-m", "bytewax.run", "wc_flow:flow"], capture_output=True, text=True, timeout=120,
                       cwd=tmp_path, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == ["('be', 2)", "('is', 1)", "('not', 1)", "('or', 1)", "('that', 1)", "('to', 2)"]


def test_alias_modules_are_the_implementation():
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    code = ("import bytewax.operators as a, bytewax_b200.operators as b, bytewax.operators.windowing as w, "
            "bytewax_b200.operators.windowing as w2, bytewax.dataflow as d, bytewax_b200.dataflow as d2; "
            "assert a.map is b.map and a.StatefulLogic is b.StatefulLogic and w is w2 and d is d2; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60, env=env)
    assert r.stdout.strip() == "ok", r.stderr
