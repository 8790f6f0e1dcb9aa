#!/usr/bin/env python
"""Example 1 of the reference (scripts/BHM-example.jl), line for line, on the B200 path.

Ground state of a 1-D Bose-Hubbard chain, 6 particles in 6 sites: FCIQMC with 1000 walkers, shift and projected energy with
blocking-analysis error bars, compared with the Lanczos ground state computed through the same device `mul!`.
Run on a machine with a GPU:   python examples/bhm_example.py
(The same calls, with Rimu's names; the Julia shim `julia/RimuB200.jl` exposes them to Rimu.jl itself.)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rimu_b200 as R  # noqa: E402

# ## Setting up the model (BHM-example.jl:20-24)
initial_address = R.near_uniform(R.BoseFS, 6, 6)
H = R.HubbardReal1D(initial_address, u=6.0, t=1.0)

# ## Parameters of the calculation (:31-46)
target_walkers = 1_000
steps_equilibrate = 1_000
steps_measure = 2_000
last_step = steps_equilibrate + steps_measure
time_step = 0.001

# ## Defining an observable (:58-64): the projected energy onto the starting vector
initial_vector = R.default_starting_vector(initial_address, style=R.IsDynamicSemistochastic())
post_step_strategy = R.ProjectedEnergy(H, initial_vector)

# ## Running the calculation (:70-82).  `solve` hands batches of 64 steps to the device (rimu_advance): the shift update, the
# abort rules and the two projections run on the GPU; the report below has one row per step all the same.
problem = R.ProjectorMonteCarloProblem(H, start_at=initial_vector, last_step=last_step, time_step=time_step,
                                       target_walkers=target_walkers, post_step_strategy=post_step_strategy)
simulation = R.solve(problem)
df = R.DataFrame(simulation)
print(df.tail(3).to_string())

# ## Analysing the results (:104-110)
se = R.shift_estimator(df, skip=steps_equilibrate)
pe = R.projected_energy(df, skip=steps_equilibrate)

# exact reference (:137-139 use ExactDiagonalizationProblem; here: Lanczos over the device mul!)
start = R.GPUDVec([(initial_address, 1.0)], style=R.IsDeterministic())
vals, vecs, info = R.eigsolve_lanczos(H, start, krylovdim=60, tol=1e-10, maxiter=20)
print(f"""
Energy from {steps_measure} steps with {target_walkers} walkers:
Shift: {se.mean} ± {se.err}
Projected Energy: {pe.f} ± {pe.sigma_f}
Exact Energy: {vals[0]}   (dimension {R.dimension(H)})
""")
assert abs(se.mean - (-4.0215)) < 0.1 * 4.0215  # the script's own check (:157-158)
