"""numpy emulation of the three-pass plan for large 2-D transforms (DESIGN section 10, fft2 8192^2), at a small scale.

R = A*LA rows, C = LB*Cb columns.  Full size: A = 512, LA = 16, LB = 32, Cb = 256.
  pass 1  length-A transforms over the high row digit (stride LA rows), twiddle W_R^(k1*r_lo) on store   [existing pass A]
  pass 2  per 16-row group k1 and column residue c_rest: the LA*LB blocks e = LB*r_lo + c_hi sit at stride Cb in the
          flattened group, so this is ONE strided (LA*LB)-point tile whose Stockham stages skip the twiddles after the
          first radix-LA stage (= the LA x LB two-dimensional transform); output q = k2 + LA*kc1 is multiplied by
          W_C^(kc1*c_rest) and stored at row k1 + A*k2, column kc1*Cb + c_rest                               [new operator]
  pass 3  length-Cb transforms of contiguous row segments, output column kc1 + LB*kc2                        [existing pass B]
Prints the maximum deviation from numpy.fft.fft2.
"""
import numpy as np

rng = np.random.default_rng(0)
LA, LB, A, Cb = 4, 8, 8, 8
R, C = A * LA, LB * Cb
X = rng.standard_normal((R, C)) + 1j * rng.standard_normal((R, C))
W = lambda n, e: np.exp(-2j * np.pi * e / n)

Y1 = np.zeros_like(X)
for r_lo in range(LA):
    F = np.fft.fft(X[r_lo::LA, :], axis=0)
    for k1 in range(A):
        Y1[k1 * LA + r_lo, :] = F[k1] * W(R, k1 * r_lo)

T = np.zeros_like(X)
flat = Y1.reshape(A, LA * C)
for k1 in range(A):
    for c_rest in range(Cb):
        d = flat[k1, c_rest::Cb].reshape(LA, LB)          # e = LB*r_lo + c_hi
        s = np.fft.fft(np.fft.fft(d, axis=0), axis=1)     # DIF stages without the inter-stage twiddles
        for q in range(LA * LB):
            k2, kc1 = q % LA, q // LA
            T[k1 + A * k2, kc1 * Cb + c_rest] = s[k2, kc1] * W(C, kc1 * c_rest)

out = np.zeros_like(X)
for r in range(R):
    for kc1 in range(LB):
        out[r, kc1 + LB * np.arange(Cb)] = np.fft.fft(T[r, kc1 * Cb:(kc1 + 1) * Cb])
print("max |three-pass - fft2| =", np.abs(out - np.fft.fft2(X)).max())
