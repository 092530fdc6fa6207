#!/bin/bash
# rebuild the CUDA library; non-zero exit (and the compiler output) on failure
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; g.build()" > /tmp/snn_build.log 2>&1 || { tail -30 /tmp/snn_build.log; exit 1; }
tail -1 /tmp/snn_build.log
