#!/bin/bash
# second round-2 session: ncu --set full pages of the two MAGVIT2 kernels that replaced round-2 ones (8-row stem conv,
# mma.sync output conv), in situ in one encode + decode pass of 32 images (scripts/magvit_one_pass.py); raw page exported
# to CSV on the box.  scripts/ncu_summary.py turns it into profiles/r02b_ncu_kernels.{md,json}.
set -u
D=gpurun_out/ncu_r2b
mkdir -p $D
timeout -k 10 100 ncu --set full --clock-control none -f --profile-from-start off -k "regex:out_conv_mma|stem_conv" -c 2 \
  -o $D/r02b_vq_new python scripts/magvit_one_pass.py 32 > $D/r02b_vq_new.log 2>&1
echo "ncu rc=$?"
ncu -i $D/r02b_vq_new.ncu-rep --page raw --csv > $D/r02b_vq_new_raw.csv 2>/dev/null
rm -f $D/r02b_vq_new.ncu-rep
gzip -f $D/r02b_vq_new_raw.csv
ls -la $D; tail -3 $D/r02b_vq_new.log
