"""Synthetic keyswitch problems shared by tests and bench."""
import numpy as np

import oracle_binding as ob


class KsProblem:
    def __init__(self, n, D, K, batch, bits=51, seed=1234, moduli=None):
        o = ob.oracle()
        self.n, self.D, self.K, self.batch = n, D, K, batch
        self.moduli = np.array(moduli if moduli is not None else ob.primes(K, bits, n), dtype=np.uint64)
        qk = int(self.moduli[K - 1])
        self.msf = np.array([o.ho_inv_mod(qk % int(q), int(q)) for q in self.moduli], dtype=np.uint64)
        self.msf[K - 1] = 0
        # keys[j][(c*K + i)*n + l] uniform in [0, q_i)
        self.keys = []
        for j in range(D):
            k = np.zeros((2, K, n), dtype=np.uint64)
            for c in range(2):
                for i in range(K):
                    k[c, i] = ob.splitmix(n, seed + 1000 * j + 10 * i + c, int(self.moduli[i]))
            self.keys.append(k.reshape(-1))
        t = np.zeros((batch, D, n), dtype=np.uint64)
        r = np.zeros((batch, 2, D, n), dtype=np.uint64)
        for b in range(batch):
            for j in range(D):
                q = int(self.moduli[j])
                t[b, j] = ob.splitmix(n, seed + 7919 * b + j + 1, q)
                r[b, 0, j] = ob.splitmix(n, seed + 7919 * b + 100 + j, q)
                r[b, 1, j] = ob.splitmix(n, seed + 7919 * b + 200 + j, q)
        self.t_target = t.reshape(batch, -1)
        self.result = r.reshape(batch, -1)

    def expected(self, alt=False):
        if alt:
            return np.stack([
                ob.keyswitch(self.result[b], self.t_target[b], self.n, self.D, self.K, self.moduli, self.keys,
                             self.msf, 1, alt=True) for b in range(self.batch)])
        return ob.keyswitch(self.result.reshape(-1), self.t_target.reshape(-1), self.n, self.D, self.K,
                            self.moduli, self.keys, self.msf, self.batch).reshape(self.batch, -1)
