set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
python profiles/ncu_targets.py 592 42 128 > gpurun_out/r2i_plain_c3.log 2>&1 || exit 1
python profiles/ncu_targets.py 256 4 1 > gpurun_out/r2i_plain_c2.log 2>&1 || exit 1
python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_plain_c3s.log 2>&1 || exit 1
# launch lists
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2i_launches_c3.csv python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu_l3.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2i_launches_c2.csv python profiles/ncu_targets.py 256 4 1 > gpurun_out/r2i_ncu_l2.log 2>&1
# full captures
FULL="$NCU --set full --import-source on --kernel-name-base demangled"
timeout 400 $FULL -k regex:decode_self_attention_tm -s 498 -c 2 -f -o gpurun_out/r2i_selfattn python profiles/ncu_targets.py 592 42 128 > gpurun_out/r2i_ncu1.log 2>&1
timeout 300 $FULL -k "regex:ffn_fused_tc<1, 0, 1>" -s 20 -c 2 -f -o gpurun_out/r2i_ffn python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu2.log 2>&1
timeout 300 $FULL -k regex:decode_cross_attention_tc -s 20 -c 2 -f -o gpurun_out/r2i_cross python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu3.log 2>&1
timeout 300 $FULL -k "regex:gemm_bf16_tc<0>" -s 260 -c 2 -f -o gpurun_out/r2i_gemm python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu4.log 2>&1
timeout 300 $FULL -k regex:sample_tokens -s 2 -c 2 -f -o gpurun_out/r2i_sample python profiles/ncu_targets.py 592 2 128 > gpurun_out/r2i_ncu5.log 2>&1
timeout 300 $FULL -k regex:decode_attn -s 30 -c 2 -f -o gpurun_out/r2i_decattn python profiles/ncu_targets.py 256 4 1 > gpurun_out/r2i_ncu6.log 2>&1
ls -la gpurun_out/r2i_* | awk '{print $5, $9}'
