#!/bin/bash
for cap in 134217728 268435456 1073741824 4294967296; do
  python tools/search_once.py s7 262144 6000000 $cap 2>&1 | tail -1 | python -c "
import sys,ast; d=ast.literal_eval(sys.stdin.read().strip()); print('cap', sys.argv[1], 'exp', d['expansions'], 'kernel_ms %.1f' % d['kernel_ms'], 'Mexp/s %.1f' % (d['expansions']/d['kernel_ms']/1e3), 'rounds', d['rounds'])" $cap
done
