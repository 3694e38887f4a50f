# eik_fine_kernel: resident warps (share of the free device memory for the per-warp windows) against speed, on the GPU box
for f in ${FRACS:-0.15 0.25 0.35 0.5 0.7 0.85}; do echo "MCMCEQ_SCRATCH_FRACTION=$f" >> gpurun_out/${TAG:-finew}.log; MCMCEQ_SCRATCH_FRACTION=$f python tools/fine_probe.py ${CH:-16} 2>&1 | tail -n 1 >> gpurun_out/${TAG:-finew}.log; done; cat gpurun_out/${TAG:-finew}.log
