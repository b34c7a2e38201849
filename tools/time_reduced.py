import os
import sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reducedbasismethods.jl_b200"))
import numpy as np, torch
import rbm_b200 as r
ctx = r.Context(0); r.set_default_context(ctx)
L = r.BumpOnTail.domain_length(0.3)
for n in (10_000, 100_000):
    nh, nt = 16, 250
    x, v, w = r.BumpOnTail.draw_accept_reject(n, seed=1234)
    particles = r.ParticleList(x[None], v[None], w[None]); poisson = r.PoissonSolverPBSplines(3, nh, L)
    train = np.linspace(1/3, 5/3, 10)
    ip = r.IntegratorParameters(0.1, nt, nt + 1, nh, n, 10)
    ts = r.TrainingSet.create(particles, poisson, nt + 1, {"kappa": 0.3}, r.ParameterSpace(r.Parameter("chi", train)), ip)
    r.integrate_vp_all(particles, poisson, train, ip, ts.snapshots)
    rb = r.ReducedBasis.from_trainingset(r.CotangentLiftEVD(), ts, particle_tol=1e-9, field_tol=1e-5)
    rfield = r.DEIMElectricField(poisson, rb.Psi_p, rb.Psi_e, rb.Pi_e)
    chis = np.sort(np.random.default_rng(1234).uniform(1/3, 5/3, 256))
    model = r.ReducedModel(particles, rfield, ctx)
    W = np.zeros((256, 2)); 
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model.run(chis, 0.1, nt, 2, W=W)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    # full model on the same sweep for comparison
    h = r.PICHandle(n, poisson, ctx); h.set_particles(x, v, float(w[0])); W2 = np.zeros((256, 2))
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h.run(chis, 0.1, nt, 2, W=W2); torch.cuda.synchronize(); dtf = time.perf_counter() - t0
    print(f"np={n} k_p={rb.k_p} k_e={rb.k_e}: reduced+DEIM 256 samples x {nt} steps: {dt:.3f} s ({256*nt/dt:.0f} sample-steps/s) | full model {dtf:.3f} s | W rel diff {np.abs(W[:,1]-W2[:,1]).max()/np.abs(W2[:,1]).max():.2e}")
    model.close(); h.close()
