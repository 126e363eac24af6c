"""``bytewax.visualize`` (pysrc/bytewax/visualize.py): the step tree of a dataflow as JSON or a Mermaid graph."""
import json
from typing import Any, Dict, List

from bytewax_b200.dataflow import Dataflow


def _steps(obj) -> List[Dict[str, Any]]:
    out = []
    for st in getattr(obj, "substeps", []):
        out.append({"typ": "RenderedOperator", "op_type": type(st).__name__, "step_name": getattr(st, "step_name", ""),
                    "step_id": st.step_id, "substeps": _steps(st)})
    return out


def to_json(flow: Dataflow) -> str:
    """Encode the dataflow's operator tree as JSON."""
    return json.dumps({"typ": "RenderedDataflow", "flow_id": flow.flow_id, "substeps": _steps(flow)}, indent=2)


def to_mermaid(flow: Dataflow) -> str:
    """A Mermaid flowchart of the top-level steps, in definition order."""
    lines = ["flowchart TD", f'subgraph "{flow.flow_id} (Dataflow)"']
    prev = None
    for st in _steps(flow):
        lines.append(f'{st["step_id"]}["{st["step_name"] or st["step_id"]} ({st["op_type"]})"]')
        if prev is not None:
            lines.append(f'{prev} --> {st["step_id"]}')
        prev = st["step_id"]
    lines.append("end")
    return "\n".join(lines)
