"""Small MJCF models used by the analytic known-answer tests (written for this repo)."""

FREE_SPHERE = """
<mujoco><compiler angle="radian"/><option timestep="0.001" gravity="0 0 -9.81" integrator="{integ}"/>
<worldbody><body name="ball" pos="0 0 5"><freejoint/><geom type="sphere" size="0.1" density="1000"/></body></worldbody>
</mujoco>"""

PENDULUM = """
<mujoco><compiler angle="radian"/><option timestep="0.0005" gravity="0 0 -9.81" integrator="RK4"/>
<worldbody><body name="arm" pos="0 0 2"><joint name="h" type="hinge" axis="0 1 0"/>
<geom type="sphere" size="0.05" pos="0 0 -1" mass="1"/></body></worldbody></mujoco>"""

# free-floating branched chain, no gravity, no damping: momentum and energy are conserved
FLOATING_CHAIN = """
<mujoco><compiler angle="radian"/><option timestep="0.0005" gravity="{g}" integrator="RK4"/>
<worldbody>
 <body name="base" pos="0 0 3" quat="0.9 0.1 0.3 0.2"><joint type="free"/>
  <geom type="box" size="0.2 0.1 0.05" density="800"/>
  <body name="l1" pos="0.2 0 0" euler="0.3 0.2 0.1"><joint name="j1" type="hinge" axis="0 1 0" pos="0 0 0"/>
   <geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.04" density="900"/>
   <body name="l2" pos="0.3 0 0"><joint name="j2" type="hinge" axis="0 0 1"/><joint name="j2s" type="slide" axis="1 0 0"/>
    <geom type="capsule" fromto="0 0 0 0.2 0.1 0" size="0.03" density="900"/>
    <geom type="sphere" pos="0.2 0.1 0" size="0.06" density="500"/>
   </body>
  </body>
  <body name="r1" pos="-0.2 0 0"><joint name="j3" type="hinge" axis="1 1 0"/>
   <geom type="cylinder" fromto="0 0 0 -0.25 0 0.1" size="0.05" density="700"/>
  </body>
 </body>
</worldbody></mujoco>"""

SPHERE_ON_PLANE = """
<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81" solver="{solver}" iterations="200" tolerance="1e-14"/>
<worldbody><geom name="floor" type="plane" size="5 5 0.1"/>
<body name="ball" pos="0 0 0.1"><freejoint/><geom name="ball" type="sphere" size="0.1" density="1000" friction="0.8 0.005 0.0001"/></body>
</worldbody></mujoco>"""

BOX_ON_PLANE = """
<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81" solver="{solver}" iterations="100"/>
<worldbody><geom name="floor" type="plane" size="5 5 0.1"/>
<body name="box" pos="0 0 0.3" euler="0.1 0.2 0.3"><freejoint/><geom name="box" type="box" size="0.1 0.15 0.2" density="500"/></body>
<body name="cap" pos="1 0 0.3" euler="0.4 1.2 0"><freejoint/><geom name="cap" type="capsule" size="0.05 0.2" density="500"/></body>
</worldbody></mujoco>"""

LIMITED_HINGE = """
<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81" solver="{solver}" iterations="100" tolerance="1e-12"/>
<worldbody><body name="arm" pos="0 0 2"><joint name="h" type="hinge" axis="0 1 0" limited="true" range="-0.5 0.5" damping="0.5"/>
<geom type="capsule" fromto="0 0 0 1 0 0" size="0.05" density="1000"/></body></worldbody>
<actuator><motor joint="h" gear="2" ctrllimited="true" ctrlrange="-1 1"/></actuator></mujoco>"""
