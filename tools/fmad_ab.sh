tools/ab_variants.sh s18 main fmad
export MCMCEQ_LIB=mcmc_eq_b200/libmcmceq_b200_fmad.so
python -m pytest tests/test_bench_scale_gpu.py tests/test_eikonal_gpu.py tests/test_forward_gpu.py -x -q 2>&1 | tail -n 3
python tools/misfit_probe.py
