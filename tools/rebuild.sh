#!/bin/bash
# rebuild libfdgraph.so in-tree (development helper)
cd "$(dirname "$0")/.." && python -c "
import importlib.util
spec=importlib.util.spec_from_file_location('b','feynmandiagram.jl_b200/_build.py'); m=importlib.util.module_from_spec(spec); spec.loader.exec_module(m); m.build(force=True, verbose='$1'=='-v')" 2>&1 | grep -E "error|warning: v|registers|spill" 
exit 0
