# Runs a DF-RAGF2(None,0) calculation on the GPU -- counterpart of the reference's examples/04-ragf2.py for the
# OptRAGF2 driver.  PySCF is not needed: a seeded model molecule provides the `hf` protocol (with PySCF installed,
# pass any converged density-fitted pyscf.scf.RHF through auxgf_b200.hf.from_pyscf instead).
from auxgf_b200.agf2 import OptRAGF2
from auxgf_b200.hf import ModelRHF

rhf = ModelRHF(nao=24, nocc=5, naux=70, seed=1).run()

gf2 = OptRAGF2(rhf, verbose=False, etol=1e-7, maxiter=20)
gf2.run()
print('OptRAGF2: converged = %s  iterations = %d  E(mp2) = %.10f  E(tot) = %.10f' %
      (gf2.converged, gf2.iteration, gf2.e_mp2, gf2.e_tot))
print('IP = %.6f  EA = %.6f  chempot = %.6f' % (gf2.ip[0], gf2.ea[0], gf2.chempot))

# One sector of the self-energy through the static build_part (the call the moment kernels sit behind):
se_occ = OptRAGF2.build_part(gf2.gf.as_occupied(), gf2.gf.as_virtual(), gf2.eri)
print('occupied-sector self-energy: %d poles' % se_occ.naux)
