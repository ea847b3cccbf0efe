#!/bin/bash
# usage: tools/ncu_kernel.sh <demangled-name regex> <launch-skip> <out-base> -- <command...>
# One `ncu --set full` capture of a single launch (recipe of /opt/skills/guides/B200_PROFILING.md); run the command
# once without ncu first.
set -e
REGEX="$1"; SKIP="$2"; OUT="$3"; shift 4
"$@" > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$REGEX" -s "$SKIP" -c 1 \
    -o "$OUT" -f "$@" > "$OUT.log" 2>&1
tail -3 "$OUT.log" | cut -c1-200
ls -la "$OUT.ncu-rep"
