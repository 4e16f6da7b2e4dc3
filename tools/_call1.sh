mkdir -p gpurun_out
{ free -g | head -2; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; } > gpurun_out/box.txt 2>&1
bash tools/sweep_variants.sh default base acq acq0 > gpurun_out/sweep1.txt 2>&1
N_CTA=441 python tools/phases.py > gpurun_out/phases1.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest1.txt 2>&1
tail -3 gpurun_out/pytest1.txt
cat gpurun_out/sweep1.txt gpurun_out/phases1.txt gpurun_out/box.txt
