# SCS-RAGF2(None,0): the spin-component-scaling factors of the reference's examples/10-scs-ragf2.py (the reference's
# OptRAGF2 accepts and ignores them, auxgf/agf2/optragf2.py:33-34; here they reach the moment kernels).
from auxgf_b200.agf2 import OptRAGF2
from auxgf_b200.hf import ModelRHF

rhf = ModelRHF(nao=24, nocc=5, naux=70, seed=1).run()

gf2 = OptRAGF2(rhf, verbose=False, os_factor=6/5, ss_factor=1/3, maxiter=20)
gf2.run()
print('SCS-OptRAGF2: converged = %s  iterations = %d  E(corr) = %.12f' % (gf2.converged, gf2.iteration, gf2.e_corr))
