"""Oracle: polar AC power-flow residual and Jacobian in complex-power form.

TEST INFRASTRUCTURE ONLY.  This is what the reference's generated ``inner_F`` / ``inner_J`` compute for
the LoopEqn polar power-flow model of SURVEY.md section 8d (generated code shown in SURVEY.md Appendix
B: module_printer.py:784-1047 for F, loop_jac.py:1865-2040 for J), written independently with the
textbook matrix formulas so that it can check both the reference's output and this repo's CUDA
printer.  Layout of unknowns and equations follows the reference's ``Address`` order:
y = [Va(nb), Vm(nb)];  F = [P_eqn over PVPQ, Q_eqn over PQ, Va[ref], Vm[ref u pv] - 1].
"""
import numpy as np
import scipy.sparse as sps


class PFOracle:
    def __init__(self, Ybus, pv, pq, ref=0):
        self.Y = sps.csr_matrix(Ybus)
        self.nb = self.Y.shape[0]
        self.pv = np.asarray(pv)
        self.pq = np.asarray(pq)
        self.pvpq = np.concatenate([self.pv, self.pq])
        self.fixed = np.concatenate([[ref], self.pv])   # Vm fixed to 1 (ref and pv are 0..npv)
        self.ref = ref
        self.n = 2 * self.nb

    def F(self, y, Pinj, Qinj):
        nb = self.nb
        Va, Vm = y[:nb], y[nb:]
        V = Vm * np.exp(1j * Va)
        S = V * np.conj(self.Y @ V)
        return np.concatenate([S.real[self.pvpq] - Pinj[self.pvpq], S.imag[self.pq] - Qinj[self.pq],
                               [Va[self.ref]], Vm[self.fixed] - 1.0])

    def J(self, y):
        nb = self.nb
        Va, Vm = y[:nb], y[nb:]
        V = Vm * np.exp(1j * Va)
        I = self.Y @ V
        dV = sps.diags(V)
        dI = sps.diags(I)
        dVn = sps.diags(V / Vm)
        dS_dVa = 1j * dV @ np.conj(dI - self.Y @ dV)
        dS_dVm = dV @ np.conj(self.Y @ dVn) + np.conj(dI) @ dVn
        dS_dVa = sps.csr_matrix(dS_dVa)
        dS_dVm = sps.csr_matrix(dS_dVm)
        top = sps.hstack([dS_dVa.real[self.pvpq], dS_dVm.real[self.pvpq]])
        mid = sps.hstack([dS_dVa.imag[self.pq], dS_dVm.imag[self.pq]])
        e_ref = sps.csr_matrix(([1.0], ([0], [self.ref])), shape=(1, 2 * nb))
        nf = len(self.fixed)
        e_vm = sps.csr_matrix((np.ones(nf), (np.arange(nf), nb + self.fixed)), shape=(nf, 2 * nb))
        return sps.vstack([top, mid, e_ref, e_vm]).tocsc()
