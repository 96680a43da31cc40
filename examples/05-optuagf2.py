# Runs a DF-UAGF2(None,0) calculation on the GPU (open-shell model molecule) -- counterpart of examples/05-uagf2.py
# for the OptUAGF2 driver.
from auxgf_b200.agf2 import OptUAGF2
from auxgf_b200.hf import ModelUHF

uhf = ModelUHF(nao=16, nalph=4, nbeta=3, naux=50, seed=2).run()

gf2 = OptUAGF2(uhf, verbose=False, maxiter=10)
gf2.run()
print('OptUAGF2: converged = %s  iterations = %d  E(mp2) = %.10f  E(tot) = %.10f' %
      (gf2.converged, gf2.iteration, gf2.e_mp2, gf2.e_tot))
print('IP = %.6f  EA = %.6f' % (gf2.ip[0], gf2.ea[0]))
