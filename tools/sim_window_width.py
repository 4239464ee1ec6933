#!/usr/bin/env python
"""DESIGN TOOLING (not product, not test): how many parser steps would a block need if a step resolved W positions?
Replays the reference's greedy parse (src/compress.rs:195-317) in plain Python on the first 64KB of corpus files and
counts, per aligned W-window, 1 step + 1 per "victim" (a probed position whose candidate was inserted inside the same
step -- the window is cut there and restarted, as k1_finish does). Round-1 result, bytes resolved per step:
  alice29 W32 30.3 / W64 52.3 / W128 74.8;  html 43.9 / 65.2 / 91.5;  urls 33.7 / 53.7 / 73.7;  kppkn 25.8 / 33.1 / 38.7
i.e. a 64-position step needs 42% fewer steps on text, a 128-position one 60% fewer."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from conftest import corpus

def parse(src):
    """greedy parse as in the reference (single block <= 64KB, n >= 17); returns list of probes [(pos, cand)] in order,
    and copies [(pos,len)]; skip-stride handled."""
    n=len(src); 
    shift=24; tsize=256
    while tsize<16384 and tsize<n: shift-=1; tsize*=2
    table=[0]*tsize
    def h(p): 
        x=int.from_bytes(src[p:p+4],'little'); return ((x*0x1E35A7BD)&0xFFFFFFFF)>>shift
    probes=[]; copies=[]
    s_limit=n-15
    s=1; next_hash=h(s)
    done=False
    while not done:
        skip=32
        cand=0
        ns=s
        while True:
            s=ns
            hh=next_hash
            ns=s+(skip>>5); skip+=(skip>>5)
            if ns>s_limit: done=True; break
            next_hash=h(ns)
            cand=table[hh]; probes.append((s,cand, skip>>5))
            table[hh]=s
            if src[s:s+4]==src[cand:cand+4]: break
        if done: break
        while True:
            base=s
            # extend
            l=4
            while s+l<n and src[s+l]==src[cand+l]: l+=1
            copies.append((base,l))
            s+=l
            if s>=s_limit: done=True; break
            # insert s-1, probe s
            table[h(s-1)]=s-1
            hh=h(s); cand=table[hh]; probes.append((s,cand,1)); table[hh]=s
            if src[s:s+4]!=src[cand:cand+4]:
                s+=1; next_hash=h(s); break
    return probes, copies

def steps(probes, copies, n, W):
    # positions visited as probes; count steps: per aligned W-window, chain of cuts
    import collections
    byw=collections.defaultdict(list)
    for (p,c,st) in probes: byw[p//W].append((p,c,st))
    total=0; serial=0
    for w,lst in byw.items():
        e=w*W  # entry
        total+=1
        for (p,c,st) in lst:
            if st>1: serial+=1
            if c>=e and c<p and c>=w*W:   # candidate inserted within this step -> victim
                total+=1; e=p
    return total, len(byw)

for name in ["alice29.txt","html","urls.10K","geo.protodata","kppkn.gtb","lcet10.txt"]:
    d=corpus(name)[:65536]
    pr,cp=parse(d)
    out=[name, 'probes',len(pr),'copies',len(cp)]
    for W in (32,64,128):
        t,nw=steps(pr,cp,len(d),W)
        out+= [f'W{W}: steps {t} (windows {nw}) bytes/step {len(d)/t:.1f}']
    print(*out)


def steps_unaligned(probes, n, W):
    """windows start exactly where the previous step stopped (no alignment): [s, s+W)"""
    total = 0
    i = 0
    P = len(probes)
    while i < P:
        s = probes[i][0]
        total += 1
        j = i
        nxt = None
        while j < P and probes[j][0] < s + W:
            p, c, st = probes[j]
            if j > i and c >= s and c < p:      # victim: cut here
                nxt = j
                break
            j += 1
        i = nxt if nxt is not None else j
    return total


if __name__ == "__main__":
    print("--- unaligned windows (bytes per step)")
    for name in ["alice29.txt", "html", "urls.10K", "kppkn.gtb"]:
        d = corpus(name)[:65536]
        pr, cp = parse(d)
        print(name, *[f"W{W}: {len(d) / steps_unaligned(pr, len(d), W):.1f}" for W in (32, 64, 128)])
