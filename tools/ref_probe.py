"""Time the unmodified reference (oracle/_ref) on a few ensemble members (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    from oracle import ref_runner
    from hydromodel_b200.synthetic import default_params, ensemble_member, synthetic_forcing
    SITE = os.path.join(ROOT, "hydromodel_b200", "model_parameters", "site_information.json")
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    print(ref_runner.importable())
    ic = np.load(os.path.join(ROOT, "hydromodel_b200", "model_parameters", "ic_well10_default.npy"))
    base = default_params(SITE)
    members = []
    for c in range(procs):
        p, m = ensemble_member(base, c, perturb_theta=True)
        members.append((p, c, m))
    t = time.time()
    pool = ref_runner.ReferencePool(ic, SITE, procs)
    pool.warm(members, synthetic_forcing(3650))
    print("warm", time.time() - t)
    wall, win, cpu, _ = pool.run(members, synthetic_forcing(3650), rows)
    print(wall, win, win / wall, "windows/s", cpu)
    pool.close()
