import sys; sys.path.insert(0, "/root/repo")
from mujoco_gymnasium_environments_b200.vector_env import B200VectorEnv
for t in ("quadruped_parkour", "humanoid_dancing", "humanoid_soccer", "bipedal_rescue", "humanoid_construction", "humanoid_martial_arts", "robotic_arm_assembly"):
    e = B200VectorEnv(t, 8); b = e.batch
    print(f"{t}: nq {b.nq} nv {b.nv} nu {b.nu} nbody {b.nbody} | epb {b.envs_per_block} smem/CTA {b.smem_bytes} ws/env {b.ws_bytes} arena {b.arena_floats} floats con_cap {b.con_cap} row_cap {b.row_cap} raw_cap {b.raw_cap} act_cap {b.act_cap} | wide: con {b.wide_con_cap} rows {b.wide_row_cap} arena {b.wide_arena_floats} floats, {b.wide_kib_per_env} KiB/env")
    e.close()
