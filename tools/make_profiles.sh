#!/bin/bash
# usage: tools/make_profiles.sh <report.ncu-rep> <profiles/prefix> "<kernel substrings...>"
# Reads an `ncu --set full --import-source on` report on the CPU box and writes the three summaries kept under profiles/:
#   <prefix>_summary.md      per-launch counters (tools/ncu_summary.py)
#   <prefix>_sass_hist.md    executed-instruction histogram by SASS opcode (tools/ncu_sass_hist.py)
#   <prefix>_source_lines.md stall samples / instructions per CUDA source line (tools/ncu_source_lines.py)
rep=$1; prefix=$2; shift 2
cd "$(dirname "$0")/.."
python tools/ncu_summary.py $rep > ${prefix}_summary.md
ncu -i $rep --page source --csv --print-source sass > /tmp/_sass.csv 2>/dev/null
{ echo "# SASS opcode histogram of \`$rep\` (executed warp instructions; ncu source page)"; python tools/ncu_sass_hist.py /tmp/_sass.csv 28; } > ${prefix}_sass_hist.md
ncu -i $rep --page source --csv --print-source cuda,sass > /tmp/_src.csv 2>/dev/null
{ echo "# Stall samples and executed instructions per CUDA source line of \`$rep\`"; echo; for k in "$@"; do echo '```'; python tools/ncu_source_lines.py /tmp/_src.csv "$k" 45 | cut -c1-170; echo '```'; echo; done; } > ${prefix}_source_lines.md
ls -la ${prefix}_*
