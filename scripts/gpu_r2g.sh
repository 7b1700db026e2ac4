#!/bin/bash
# experiment: does the row pitch / the relative placement of the 4 result arrays matter for the C2 kernel?
mkdir -p gpurun_out; O=gpurun_out/r2g_pitch_experiment.txt; : > $O
for pad in 0 128 384 1152 4224 16512 65664 1048704; do python scripts/sweep.py --system double_gauss --pad $pad default >> $O 2>&1; done
for gap in 4096 65536 1048576 2097152 16781312 100663808; do python scripts/sweep.py --system double_gauss --gap $gap default >> $O 2>&1; done
python scripts/sweep.py --system zoom --pad 4224 --gap 2097152 default >> $O 2>&1
python scripts/sweep.py --system zoom default >> $O 2>&1
grep -v "^$" $O
