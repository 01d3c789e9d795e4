import numpy as np
from scipy import signal
FS=2.4e6; FS1=240e3; FS2=120e3
sos = signal.cheby1(8, 0.05, 0.08, output='sos')
bb, ab = signal.butter(4, 12500/120000)
def C2(f):
    w, h = signal.sosfreqz(sos, worN=2*np.pi*np.asarray(f)/FS)
    return np.abs(h)**2
def B2(f):
    w, h = signal.freqz(bb, ab, worN=2*np.pi*np.asarray(f)/FS1)
    return np.abs(h)**2
def symresp(p_half, f, fs):
    # p_half[0] center, p_half[k] = tap +-k
    k = np.arange(1, len(p_half))
    return p_half[0] + 2*np.cos(2*np.pi*np.outer(f, k)/fs) @ p_half[1:]

def design_proto(H, fmax=60e3, iters=6, ngrid=241):
    f = np.linspace(-fmax, fmax, ngrid)
    k = np.arange(0, H+1)
    def cmat(ff):
        c = 2*np.cos(2*np.pi*np.outer(ff, k)/FS); c[:,0]=1; return c
    c0 = cmat(np.array([0.0]))[0]
    P_in = np.ones_like(f)
    Cb = C2(f)*B2(f)
    for it in range(iters):
        W = Cb/np.maximum(np.abs(P_in), 1e-3)
        Q = np.zeros((H+1,H+1))
        for m in range(1, 6):
            for sgn in (1,-1):
                if m==5 and sgn==-1: continue  # 1.2MHz +-: f+1200 and f-1200 alias same set
                A = cmat(f + sgn*m*FS1) * W[:,None]
                Q += A.T@A
        Q += 1e-18*np.trace(Q)/len(Q)*np.eye(H+1)
        sol = np.linalg.solve(Q, c0)
        p = sol/(c0@sol)
        P_in = cmat(f)@p
    # evaluate leak
    leak = 0
    for m in range(1,6):
        for sgn in (1,-1):
            leak = np.maximum(leak, np.abs(cmat(f+sgn*m*FS1)@p)*Cb/np.abs(P_in))
    return p, f, P_in, leak

if __name__ == "__main__":
    for H in (20, 24, 27, 30, 34, 40, 45, 50, 60):
        p, f, P_in, leak = design_proto(H)
        print(H, 2*H+1, "max leak %.2e" % leak.max(), "minP in |f|<50k %.3f" % np.abs(P_in[np.abs(f)<50e3]).min(), "sum|p| %.2f"%(np.abs(p[0])+2*np.abs(p[1:]).sum()))
