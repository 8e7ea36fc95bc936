import random
C=[17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]
M=(1<<32)-1
def direct(s): return [sum(C[i]*s[(i+r)%12] for i in range(12)) for r in range(12)]
def fftmds(s):
    w=lambda v: v & M
    u0=[0]*3;u2=[0]*3;ur=[0]*3;ui=[0]*3
    for b in range(3):
        e=w(s[b]+s[6+b]); o=w(s[3+b]+s[9+b])
        u0[b]=w(e+o); u2[b]=w(e-o); ur[b]=w(s[b]-s[6+b]); ui[b]=w(s[3+b]-s[9+b])
    # f0: cyclic, K=[16,32,16] (x16 deferred: a0 = Z0/16)
    t=w(u0[0]+u0[1]+u0[2])
    # Z[0]=S0K0+S1K2+S2K1 = 16 S0 +16 S1 + 32 S2 ; Z[1]=S0K1+S1K0+S2K2 = 32S0+16S1+16S2 ; Z[2]=S0K2+S1K1+S2K0=16S0+32S1+16S2
    a0=[w(t+u0[2]), w(t+u0[0]), w(t+u0[1])]
    # f2: negacyclic w=-1, K=[-1,-8,2]
    # Z[0]=S0K0 - (S1K2+S2K1) = -S0 -2S1 +8S2 ; Z[1]=S0K1+S1K0 - S2K2 = -8S0 - S1 -2S2 ; Z[2]=S0K2+S1K1+S2K0 = 2S0 -8S1 - S2
    a2=[w(8*u2[2]-u2[0]-2*u2[1]), w(-8*u2[0]-u2[1]-2*u2[2]), w(2*u2[0]-8*u2[1]-u2[2])]
    # f1: w=i, K=[(2,1),(-4,-1),(16,-1)]
    def cm(k,b): return (k[0]*ur[b]-k[1]*ui[b], k[0]*ui[b]+k[1]*ur[b])
    K1=[(2,1),(-4,-1),(16,-1)]
    def imul(z): return (-z[1], z[0])
    def add(*zs): return (w(sum(z[0] for z in zs)), w(sum(z[1] for z in zs)))
    z0=add(cm(K1[0],0), imul(cm(K1[2],1)), imul(cm(K1[1],2)))
    z1=add(cm(K1[1],0), cm(K1[0],1), imul(cm(K1[2],2)))
    z2=add(cm(K1[2],0), cm(K1[1],1), cm(K1[0],2))
    a1=[z0,z1,z2]
    out=[0]*12
    for b in range(3):
        p=w(16*a0[b]+a2[b]); q=w(16*a0[b]-a2[b])
        out[b]=w(p+a1[b][0]); out[3+b]=w(q+a1[b][1]); out[6+b]=w(p-a1[b][0]); out[9+b]=w(q-a1[b][1])
    return out
for _ in range(1000):
    s=[random.randrange(1<<22) for _ in range(12)]
    assert direct(s)==fftmds(s), (direct(s), fftmds(s))
s=[(1<<22)-1]*12
assert direct(s)==fftmds(s)
print("ok")
