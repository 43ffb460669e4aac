#!/bin/bash
# Commands behind the files in profiles/ (round 1e, round 2 at the end).  Each block is one `gpurun -- '<command>'` call on a B200 box;
# outputs land in gpurun_out/ and the summaries were copied to profiles/ by hand (see profiles/README.md).
set -e
mkdir -p gpurun_out

# tests + smoke
python -m pytest tests -m gpu -x -q
python -c "import __graft_entry__ as g; g.smoke()"

# bench lines (N = 1; N = 2 / 8 under torchrun) and the CPU arm
python bench.py > gpurun_out/bench_default.json
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_reference.json
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8

# launch list of the bench command (share of each kernel; times are cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null

# per-workload numbers and the batch-size sweep
python tools/profile_workloads.py ur10:1024 ur10:4096 ur10:16384 ur10:65536 kuka:1024 kuka:4096 kuka:16384 kuka:65536 \
    lwa4d:4096 chain20:8192 chain20:65536 kuka_table:16384 > gpurun_out/workloads.jsonl

# ncu --set full of the three trust-region kernels (bounded maxiter so that the ~40 replays finish)
ncu --set full --clock-control none --import-source on -k regex:k_rtr_fast -s 1 -c 1 -o gpurun_out/fast_tp -f \
    python tools/profile_workloads.py ur10:65536:latency:1:300 > /dev/null
ncu --set full --clock-control none --import-source on -k regex:k_rtr_fast2 -s 1 -c 1 -o gpurun_out/fast2 -f \
    python tools/profile_workloads.py chain20:16384:auto:1:60 > /dev/null
ncu --set full --clock-control none --import-source on -k regex:k_rtr_cta -s 1 -c 1 -o gpurun_out/cta -f \
    python tools/profile_workloads.py kuka_table:296:dense:1:40 > /dev/null
# read back with: ncu -i gpurun_out/<name>.ncu-rep --page raw --csv | grep -E 'pipe_fp64|issue_active|dram__bytes|registers'

# ---------------------------------------------------------------- round 2
# the driver's bench command, its reference arm, and 2 GPUs under torchrun (weak scaling; configs[3] strong scaling)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench.json
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2z_bench_ref.json
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --robot chain20 --global-batch 65536

# bound smoothing + initialisation: times, phase cycle counts (-DGIK_BI_PROFILE variant, built by --build-profile), ncu
python tools/bi_bench.py ur10:65536 kuka:65536 chain20:65536 kuka_table:2048 > gpurun_out/r2q_bi_bench.jsonl
python tools/bi_bench.py --build-profile && python tools/bi_bench.py ur10:4 chain20:4 kuka_table:4 --profile > gpurun_out/r2q_bi_profile.log
ncu --set full --import-source on --clock-control none -k regex:k_bounds_init --launch-skip 2 -c 1 -o gpurun_out/r2q_bi_chain20 -f \
    python tools/bi_bench.py chain20:65536 > /dev/null
python tools/ncu_summary.py gpurun_out/r2q_bi_chain20.ncu-rep > profiles/r2q_k_bounds_init_chain20_b65536.txt

# dense kernel: iteration rate (A/B against another build with --lib=<name>.so in graphik_b200/lib/) and ncu with source
python tools/profile_workloads.py kuka_table:2368:dense:2:60
ncu --set full --import-source on --clock-control none -k regex:k_rtr_cta --launch-skip 1 -c 1 -o gpurun_out/r2r_cta -f \
    python tools/profile_workloads.py kuka_table:296:dense:1:40 > /dev/null
# two-problems-per-warp kernel in the throughput regime
python tools/profile_workloads.py ur10:65536:throughput:2:300

# sanitizers over every kernel variant (incl. park / resume, N = 166 fallback)
compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_smoke.py
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py

# ---- round 2, second part: conjugate-gradient solve and CIDGIK (r2ac .. r2aj)
python -m pytest tests -m gpu -q                                   # gpurun_out/r2al_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2al_bench.json # configs block now ends with BASELINE configs[4]
for b in 1024 4096 16384 65536; do python tools/cidgik_bench.py --batch $b --reps 5 --warmup 2; done > gpurun_out/r2al_cidgik_batches.jsonl
python tools/cidgik_bench.py --robot kuka --batch 1024 --reps 5 --warmup 2 >> gpurun_out/r2al_cidgik_batches.jsonl
python tools/cidgik_bench.py --robot lwa4d --batch 1024 --reps 5 --warmup 2 >> gpurun_out/r2al_cidgik_batches.jsonl
ncu --set full --clock-control none --import-source on -k regex:k_sdp -c 1 -f -o gpurun_out/r2al_k_sdp \
    python tools/cidgik_bench.py --batch 16384 --reps 1 --warmup 0 > /dev/null
python tools/ncu_summary.py gpurun_out/r2al_k_sdp.ncu-rep > profiles/r2al_k_cidgik_fused_ur10_b16384.txt
# (earlier kernel versions: r2ae = 64-thread CTAs, r2ag = + one-slot Cholesky; same tool with --batch 1024, and
#  ncu --metrics gpu__time_duration.sum ... --csv --log-file gpurun_out/r2ae_cidgik_launches.csv for the launch list)
compute-sanitizer --tool memcheck python tools/sanitize_smoke.py  > gpurun_out/r2al_memcheck.log
compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_smoke.py > gpurun_out/r2al_racecheck.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ak_cidgik_launches.csv \
    python tools/cidgik_bench.py --reps 1 --warmup 1 > /dev/null
# (r2aj_*: the same commands at the commit before the fused launch, i.e. one k_sdp + one k_fantope launch per convex iteration)
