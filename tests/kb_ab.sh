#!/bin/bash
# Developer A/B helper (test infrastructure): kernel-only times of several builds of the CUDA library, interleaved.
#   tests/kb_ab.sh [-p precision] [-c columns] [-n rounds] lib1.so lib2.so ...   ("default" = the in-tree library)
prec=double; cols=65536; rounds=2
while getopts "p:c:n:" o; do case $o in p) prec=$OPTARG;; c) cols=$OPTARG;; n) rounds=$OPTARG;; esac; done
shift $((OPTIND-1))
for i in $(seq $rounds); do
  for lib in "$@"; do
    if [ "$lib" = default ]; then unset CS2_LIB; else export CS2_LIB=$PWD/$lib; fi
    python tests/kbench.py --columns $cols --precision $prec --reps 30 --no-parity 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-44s' % '$lib', ' '.join('%s %.4f' % (k, d[k]['ms']) for k in ('sat','nl','tl','ad','ad_ckpt')))"
  done
done
