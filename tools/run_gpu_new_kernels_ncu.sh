#!/bin/bash
# ncu --set full of the kernels added in the third session (QR panel / split-row update, complex leaf panel / packs).
TAG=${1:-r1c}
mkdir -p gpurun_out
cap() {  # name regex skip what n
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_$1_${TAG} -f python tools/ncu_one.py $4 $5 > gpurun_out/ncu_$1_${TAG}.log 2>&1
}
cap qr_panel qr_panel_kernel 8 dgeqrf 8192
cap qr_vtc qr_vtc_kernel 8 dgeqrf 8192
cap qr_trmm qr_trmm_tt_kernel 8 dgeqrf 8192
cap cx_leaf panel_cx_kernel 16 zgetrf 8192
cap cx_pack_b pack_b_kernel 40 zgetrf 8192
python tools/ncu_summary.py gpurun_out/prof_qr_panel_${TAG}.ncu-rep gpurun_out/prof_qr_vtc_${TAG}.ncu-rep gpurun_out/prof_qr_trmm_${TAG}.ncu-rep gpurun_out/prof_cx_leaf_${TAG}.ncu-rep gpurun_out/prof_cx_pack_b_${TAG}.ncu-rep > gpurun_out/ncu_new_kernels_${TAG}.txt 2>&1
cat gpurun_out/ncu_new_kernels_${TAG}.txt
