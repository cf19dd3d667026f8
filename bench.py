"""bench.py -- env-steps/s of the batched env.step hot path (BASELINE.json metric) on N B200s of one node.

python bench.py --gpus N --steps K --warmup W           this repo's CUDA path (quadruped_parkour, 4096 envs/GPU)
python bench.py --task humanoid_dancing ...              BASELINE.json configs[2] (8192 envs/GPU, RK4, self contacts)
python bench.py --impl reference ...                     the CPU arm: the reference's algorithm restated (oracle port),
                                                         one process per env on all host cores
One "step" = one env.step over the whole batch (clip, frame_skip x mj_step, obs, reward, termination, same-step auto-reset).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (step+reward+obs)"
UNIT = "env-steps/s"
# --task selects the workload; the default is BASELINE.json configs[1] (the config the metric is quoted on at N=1)
WORKLOADS = {
    "quadruped_parkour": (4096, 6000, "quadruped_parkour_env: {n} envs/GPU lockstep, frame_skip 10 (dt 1 ms), Euler, PGS-50, 48 plane "
                          "contact pairs, uniform random actions over action_space x {s}, same-step auto-reset"),
    "humanoid_dancing": (8192, 1500, "humanoid_dancing_env: {n} envs/GPU lockstep, 1 RK4 step (dt 16.67 ms, 4 forward passes), PGS-50, "
                         "106 self-contact candidate pairs, uniform random actions over action_space x {s}, same-step auto-reset"),
    "humanoid_soccer": (4096, 1500, "humanoid_soccer_env: {n} envs/GPU lockstep, 1 Euler step (dt 20 ms), PGS-50, free ball + box field, "
                        "251 candidate pairs, uniform random actions over action_space x {s}, same-step auto-reset"),
    "bipedal_rescue": (2048, 300, "bipedal_rescue_env: {n} envs/GPU lockstep, 1 RK4 step (dt 20 ms, 4 forward passes), PGS-50, 63 dofs, "
                       "3175 candidate pairs, uniform random actions over action_space x {s}, same-step auto-reset"),
    "humanoid_construction": (2048, 300, "humanoid_construction_env: {n} envs/GPU lockstep, 1 RK4 step (dt 2 ms, 4 forward passes), Newton-100, "
                              "99 dofs in 12 trees, 1202 candidate pairs, uniform random actions over action_space x {s}, same-step auto-reset"),
    "humanoid_martial_arts": (4096, 1500, "humanoid_martial_arts_env: {n} envs/GPU lockstep, 1 Euler step (dt 16.67 ms), Newton-50, 47 dofs in 4 trees "
                              "(free humanoid, two free cylinder dummies, hinged board), 294 candidate pairs, uniform random actions over action_space x {s}, "
                              "same-step auto-reset"),
    "robotic_arm_assembly": (2048, 100, "robotic_arm_assembly_env: {n} envs/GPU lockstep, frame_skip 10 (dt 2 ms), Euler, Newton-50, 63 dofs in 10 trees "
                             "(7-dof arm + gripper, nine free components), 785 candidate pairs incl. 18 condim-6, uniform random actions over "
                             "action_space x {s}, same-step auto-reset"),
}
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel at the task's default size, from the
# `ncu --set full` captures summarised under profiles/ (r01_i quadruped, r01_f dancing, r01_g soccer, r01_h rescue,
# r01_j construction, r01_k martial arts, r01_l arm).
# Below the algorithmic bytes because the state written by the previous launch is still in the 126 MB L2.
NCU_TRAFFIC = {"quadruped_parkour": (4096, 4.233728e6 + 287.744e3), "humanoid_dancing": (8192, 9.365248e6 + 721.408e3),
               "humanoid_soccer": (4096, 4.960768e6 + 33.024e3), "bipedal_rescue": (2048, 3.231744e6 + 72.96e3),
               "humanoid_construction": (2048, 4.57088e6 + 48.384e3), "humanoid_martial_arts": (4096, 5.356544e6 + 18.688e3),
               "robotic_arm_assembly": (2048, 3.184896e6 + 268.032e3)}
TASK = "quadruped_parkour"
WORKLOAD = WORKLOADS[TASK][2]


def cpu_arm(n_steps, action_scale):
    from oracle import cpu_bench
    cores = len(os.sched_getaffinity(0))
    total, wall, inner = cpu_bench.run(TASK, cores, n_steps, action_scale)
    return dict(value=total / inner, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} processes x {n_steps} control steps of the fp64 oracle port (oracle/mjstep_ref.c + "
                       f"oracle/tasks_ref.py), {inner:.1f} s")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.lines = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True); self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def _finite(x):
    """JSON has no Infinity / NaN (rescue pays +inf on the first step of every episode): non-finite floats become null."""
    if isinstance(x, dict):
        return {k: _finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_finite(v) for v in x]
    if isinstance(x, float) and (x != x or x in (float("inf"), float("-inf"))):
        return None
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--task", default="quadruped_parkour", choices=sorted(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="0 = the task's BASELINE.json size")
    ap.add_argument("--action-scale", type=float, default=1.0)
    ap.add_argument("--cpu-steps", type=int, default=0, help="control steps per CPU process (0 = sized for ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    global TASK, WORKLOAD
    TASK = a.task; WORKLOAD = WORKLOADS[TASK][2]
    a.envs_per_gpu = a.envs_per_gpu or WORKLOADS[TASK][0]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3)

    if a.impl == "reference":
        if rank != 0:
            return
        # each "step" is a bounded sample: every host core steps its env `per` times
        per = a.cpu_steps or 40
        from oracle import cpu_bench
        cores = len(os.sched_getaffinity(0))
        cpu_bench.run(TASK, cores, max(3, min(per, 5 * warmup)), a.action_scale)
        total, wall, inner = cpu_bench.run(TASK, cores, per * a.steps, a.action_scale)
        v = total / inner
        print(json.dumps(dict(
            metric=METRIC, value=v, unit=UNIT, impl="reference", n_gpus=a.gpus, steps=a.steps, warmup=warmup,
            ms_per_step=inner / a.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic", config=dict(workload=WORKLOAD.format(n=a.envs_per_gpu, s=a.action_scale), envs_per_step=cores * per),
            cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind="port",
                              sample=f"{cores} processes x {per * a.steps} control steps, one env per process (oracle port; "
                                     f"mujoco/gymnasium are not installable here)"),
            e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)))
        return

    # ---- CPU baseline first (spawned processes; before CUDA is initialised in this one)
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_arm(a.cpu_steps or WORKLOADS[TASK][1], a.action_scale)

    import numpy as np
    import torch
    import torch.distributed as dist
    from mujoco_gymnasium_environments_b200 import capi, sharding
    from mujoco_gymnasium_environments_b200.tasks import TASKS
    from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = TASKS[TASK]
    N = a.envs_per_gpu
    off, _ = sharding.shard_range(rank, world, N)
    env = B200VectorEnv(TASK, N, device=local, seed=1234, env_offset=off)
    env.reset()
    hi = torch.tensor(env.single_action_space.high, device=dev) * a.action_scale
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    K = a.steps
    acts = (torch.rand((warmup + K, N, spec.act_dim), device=dev, generator=gen) * 2 - 1) * hi
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        flush.fill_(0.0); env.step(acts[i])
    barrier()
    sampler = ClockSampler(local); sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches0 = capi.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        flush.fill_(float(i))                     # L2 flush between timed iterations (inside the timed region)
        ev[i][0].record(); env.batch.step(acts[warmup + i], env._obs, env._rew, env._term, env._trunc, env._final_obs); ev[i][1].record()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = capi.launch_count() - launches0
    ms = sharding.max_over_ranks(e0.elapsed_time(e1), dev)
    kern_ms = sum(s.elapsed_time(e) for s, e in ev) / K
    value = world * N * K / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry point (pinned staging, H2D + D2H inside the timed region)
    ha = acts[warmup:].cpu().numpy()
    ho = np.zeros((N, spec.obs_dim), np.float32); hr = np.zeros(N, np.float32); ht = np.zeros(N, np.uint8); hu = np.zeros(N, np.uint8)
    for i in range(3):
        env.batch.step_host(ha[i % K], ho, hr, ht, hu)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        env.batch.step_host(ha[i], ho, hr, ht, hu)
    torch.cuda.synchronize()
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    e2e = dict(value=world * N * K / e2e_s, unit=UNIT, h2d_bytes_per_step=N * spec.act_dim * 4,
               d2h_bytes_per_step=N * (spec.obs_dim * 4 + 4 + 2))

    stats = env.batch.stats()
    sharding.all_reduce_stats(stats)                      # the one collective of the path (NCCL when world > 1)
    st = sharding.stats_dict(stats.cpu().numpy())
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bytes_launch = spec.bytes_per_env_step * N
        achieved = bytes_launch / (kern_ms * 1e-3) / 1e9
        out = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=warmup, ms_per_step=ms / K,
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=WORKLOAD.format(n=N, s=a.action_scale), envs_per_gpu=N, action_scale=a.action_scale,
                        l2="256 MiB fill between timed steps, inside the timed region",
                        envs_per_cta=env.batch.envs_per_block, smem_bytes_per_cta=env.batch.smem_bytes),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=(NCU_TRAFFIC[TASK][1] if NCU_TRAFFIC.get(TASK, (0, 0))[0] == N else None), peak_source="MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                          bytes_per_env_step=spec.bytes_per_env_step, kernel_ms=kern_ms,
                          traffic_source="ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch, profiles/r01_[f-l]_*.txt",
                          note="fp32-latency bound physics: state stays on chip, HBM sees only state load/store"),
            cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks,
            episode_stats={k: st[k] for k in ("episodes", "mean_return", "mean_length", "nan_resets", "contacts_dropped",
                                              "rows_dropped", "arena_overflows", "solver_iters", "substeps", "wide_passes")})
        print(json.dumps(_finite(out)))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
