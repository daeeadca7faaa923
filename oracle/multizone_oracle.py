"""TEST INFRASTRUCTURE ONLY: NumPy generalisation of the reference's statistics and filter design (S4-S6,
``Python/apvast.py:329-414``) to Z > 2 zones, as SURVEY.md section 8(d) cfg-5 states it:
    R_B(z) = R_{z->z},  R_D(z) = sum_{z' != z} R_{z->z'},  r_B(z) = r_{z->z},
    (U, lambda) = jdiag(R_B, R_D),  w[v] = sum_{i <= v} (u_i . r_B) / (lambda_i + mu) u_i.
Every R_{z->z'} comes from the pinned two-zone oracle run on the PAIR (z, z') (its R_A_to_A and R_A_to_B), so this
file adds only the sum over zones and is validated against the two-zone oracle at Z = 2 (tests).  It does not use the
microphone-union composition of ap_vast_unofficial_b200/zones.py: the two routes are independent."""
import numpy as np

from .apvast_oracle import ApvastOracle, jdiag


class MultiZoneOracle:
    def __init__(self, block_size, rirs, filter_length, modeling_delay, reference_indices, number_of_eigenvectors, mu,
                 statistics_buffer_length, hop_size=None, seed=0, perceptual=False):
        # perceptual=True: the pair oracle (z, q) weights zone-z microphones from target z and zone-q microphones from
        # target q (apvast.py:259-262,318-319) -- every microphone is weighted from its own zone's target
        self.Z = len(rirs)
        self.V, self.mu = int(number_of_eigenvectors), float(mu)
        self.pairs = {}
        for z in range(self.Z):
            for q in range(self.Z):
                if q == z:
                    continue
                np.random.seed(seed)
                self.pairs[(z, q)] = ApvastOracle(block_size, rirs[z], rirs[q], filter_length, modeling_delay,
                                                  reference_indices[z], reference_indices[q], 1, mu,
                                                  statistics_buffer_length, hop_size, perceptual=perceptual)

    def process_input_buffers(self, inputs):
        for (z, q), o in self.pairs.items():
            o.process_input_buffers(inputs[z], inputs[q])
        self.w, self.lam, self.R_B, self.R_D, self.r_B = [], [], [], [], []
        for z in range(self.Z):
            others = [q for q in range(self.Z) if q != z]
            first = self.pairs[(z, others[0])]
            RB = np.array(first.R_A_to_A)
            rB = np.array(first.r_A)[:, 0]
            RD = sum(np.array(self.pairs[(z, q)].R_A_to_B) for q in others)
            U, D = jdiag(RB, RD)
            lam = np.diag(D)
            a = (U[:, :self.V].T @ rB) / (lam[:self.V] + self.mu)
            self.w.append(np.cumsum(U[:, :self.V] * a[None, :], axis=1).T)      # (V, n)
            self.lam.append(lam)
            self.R_B.append(RB); self.R_D.append(RD); self.r_B.append(rB)
