import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from hydromodel_b200.src.simulation import Simulation
from hydromodel_b200.synthetic import synthetic_forcing, default_params
from helpers import SITE, tol_scaled
g=np.load('/root/repo/tests/golden/windows_default.npz')
sim=Simulation('/tmp/x',seed=1234); sim.setupModel(default_params(SITE), synthetic_forcing(365))
print('ic err', tol_scaled(sim.mData['initial_cond'], g['ic']))
sim.run()
o=sim.output
rel=np.abs(o['theta_vol'][:151]-g['theta_vol'])/g['theta_vol']
print('theta rel by window (max):', np.round(rel.max(axis=1)[::10],7))
print('quantiles', np.quantile(rel,[.5,.9,.99,.999,1]))
print('wtd same', np.mean(o['wtd_est_cm'][:151]==g['wtd_est_cm']))
pe=[tol_scaled(o['psi_press'][i], g['psi_press'][i]) for i in range(0,151,10)]
print('psi tol units', np.round(pe,3))
