"""Run the reference's OWN models/graph_gen.py (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so this
module is only used by tools/make_golden.py (fixture generation) and by CPU
tests that are skipped when the reference tree is absent.  graph_gen.py imports
``open3d`` and ``tensorflow`` at module top (graph_gen.py:8-9) but uses neither
in the radius-graph builder; empty stub modules make the import succeed.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'models', 'graph_gen.py'))


def load():
    for name in ('open3d', 'tensorflow'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location(
        '_reference_graph_gen', os.path.join(REFERENCE_ROOT, 'models', 'graph_gen.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
