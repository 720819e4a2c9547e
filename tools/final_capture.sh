# ncu --set full captures of the final build (one GPU): cooperative / team-job schedule at N = 256 and N = 4096, the tcgen05
# STFT GEMM; then summarise here with tools/ncu_summary.py (see profiles/README.md).   bash tools/final_capture.sh <tag>
TAG=${1:-r02z}
NCU="ncu --set full --clock-control none"
timeout 300 $NCU -k regex:vr_fused_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/${TAG}_coop_n256 python tools/ncu_target.py 256 0 0 > gpurun_out/${TAG}_cap.log 2>&1
timeout 300 $NCU -k regex:vr_team_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/${TAG}_team_n256_overlap python tools/ncu_target.py 256 1 1 >> gpurun_out/${TAG}_cap.log 2>&1
timeout 300 $NCU -k regex:vr_team_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/${TAG}_team_n4096 python tools/ncu_target.py 4096 1 0 >> gpurun_out/${TAG}_cap.log 2>&1
timeout 300 $NCU -k regex:vr_fused_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/${TAG}_coop_n4096 python tools/ncu_target.py 4096 0 0 >> gpurun_out/${TAG}_cap.log 2>&1
timeout 300 $NCU -k regex:gemm_tf32x3 --launch-skip 1 --launch-count 1 -f -o gpurun_out/${TAG}_stft_gemm_n256 python tools/ncu_stft_target.py >> gpurun_out/${TAG}_cap.log 2>&1
timeout 300 $NCU -k regex:gemm_tf32x3 --launch-skip 3 --launch-count 1 -f -o gpurun_out/${TAG}_stft_gemm_n4096 python tools/ncu_stft_target.py >> gpurun_out/${TAG}_cap.log 2>&1
grep -c "Report:" gpurun_out/${TAG}_cap.log
ls -la gpurun_out/${TAG}_*.ncu-rep
