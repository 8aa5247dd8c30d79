"""Shared-memory bank conflicts of a 32-lane byte gather from a rotated window, by lane arrangement (pu x pv pixels of
the crop per warp load) and row pitch in 32-bit words: average and worst wavefronts per load over all angles.
Result that shaped the dense kernels: 32 x 1 lanes conflict ~2-way at every pitch; 8 x 4 lanes are conflict-free at
pitch = 2 mod 4 words and 1.17-way at 4 * odd words.  (CPU only.)"""
import numpy as np
def wavefronts(P_words, theta, rng, pu, pv, trials=60):
    tot=0
    c,s=np.cos(theta),np.sin(theta)
    for _ in range(trials):
        x0=rng.uniform(60,80); y0=rng.uniform(60,80)
        i=np.arange(32); u=i%pu; v=i//pu
        X=np.floor(x0+u*c-v*s+0.5).astype(int); Y=np.floor(y0+u*s+v*c+0.5).astype(int)
        word=Y*P_words+X//4
        bank=word%32
        m=0
        for b in np.unique(bank):
            m=max(m,len(np.unique(word[bank==b])))
        tot+=m
    return tot/trials
rng=np.random.default_rng(0)
thetas=np.linspace(0,2*np.pi,73)
for (pu,pv) in ((32,1),(16,2),(8,4),(4,8)):
    for P in (44,45,46,47,48,49,50,52,36,40,56,60,64):
        w=[wavefronts(P,t,rng,pu,pv) for t in thetas]
        print((pu,pv),P,"mean %.2f max %.2f"%(np.mean(w),np.max(w)))
