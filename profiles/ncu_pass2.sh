set -x
cd $GRAFT_REPO_ROOT
FULL="ncu --clock-control none --set full --import-source on --kernel-name-base demangled"
timeout 300 $FULL -k 'regex:ffn_fused_tc<\(int\)1, \(int\)0, \(int\)1>' -s 6 -c 2 -f -o gpurun_out/r2i_ffn python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu2.log 2>&1
timeout 300 $FULL -k 'regex:gemm_bf16_tc<\(int\)0>' -s 126 -c 2 -f -o gpurun_out/r2i_gemm python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu4.log 2>&1
timeout 300 $FULL -k 'regex:gemm_bf16_tc<\(int\)1>' -s 140 -c 1 -f -o gpurun_out/r2i_gemm_ln python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu7.log 2>&1
ls -la gpurun_out/r2i_ffn* gpurun_out/r2i_gemm*
