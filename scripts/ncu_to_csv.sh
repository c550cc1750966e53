#!/bin/bash
# usage: scripts/ncu_to_csv.sh <report.ncu-rep> <summary.csv>   (runs on the GPU box right after a capture:
# the raw report is too large to bring back, the per-kernel summary is what profiles/ keeps)
ncu -i "$1" --page raw --csv > "$1.raw.csv" 2>/dev/null && python scripts/ncu_summary.py "$1.raw.csv" "$2"
gzip -9 -c "$1.raw.csv" > "$2.raw.csv.gz"
rm -f "$1" "$1.raw.csv"
