"""Discrete simulation of the paged decode kernel ring protocol (one CTA: producer + consumer warps, mbarrier parity\nsemantics, TMA tiles landing out of order).  usage: python tools/paged_ring_sim.py NS tiles_per_item,... [max_landing_delay]\nNS=10 (not a multiple of the 4 consumer warps) fails the content assert; NS=8/12 pass."""
# discrete simulation of the paged kernel's barrier protocol (one CTA), parity semantics included
import random, sys
NS, NW = int(sys.argv[1]), 4
items = [int(x) for x in sys.argv[2].split(',')]   # tiles per item for this CTA
class Bar:
    def __init__(s, count): s.count=count; s.pending=count; s.phase=0
    def arrive(s):
        s.pending-=1
        if s.pending==0: s.phase+=1; s.pending=s.count
    def test(s, parity): return (s.phase & 1) != parity   # try_wait.parity: true iff current phase parity != parity
full=[Bar(1) for _ in range(NS)]; empty=[Bar(1) for _ in range(NS)]
wkfull=[Bar(1) for _ in range(4)]; wkempty=[Bar(NW) for _ in range(4)]
slots=[None]*4
stage_content=[None]*NS
inflight=[]  # (tile g, stage) landing later
def producer():
    n=0; it=0
    for idx,nt in enumerate(items+[None]):
        slot=it&3
        if it>=4:
            while not wkempty[slot].test(((it>>2)-1)&1): yield
        slots[slot]=(idx,nt)
        wkfull[slot].arrive()
        if nt is None: return
        it+=1
        for i in range(nt):
            g=n+i; s=g%NS
            if g>=NS:
                while not empty[s].test(((g//NS)-1)&1): yield
            inflight.append([g,s,random.randint(1,int(sys.argv[3]) if len(sys.argv)>3 else 6)])
            yield
        n+=nt
def consumer(w):
    n=0; it=0
    while True:
        slot=it&3
        while not wkfull[slot].test((it>>2)&1): yield
        idx,nt=slots[slot]
        if nt is None: return
        wkempty[slot].arrive(); it+=1
        i=(w+NW-(n%NW))%NW
        while i<nt:
            g=n+i; s=g%NS
            while not full[s].test((g//NS)&1): yield
            assert stage_content[s]==g, (w,g,s,stage_content[s])
            for _ in range(random.randint(0,3)): yield
            empty[s].arrive()
            i+=NW
        n+=nt
        for _ in range(random.randint(0,8)): yield
procs=[producer()]+[consumer(w) for w in range(NW)]
alive=[True]*5
steps=0
while any(alive):
    steps+=1
    if steps>2_000_000: print("DEADLOCK/too long"); break
    # land TMAs
    for t in inflight[:]:
        t[2]-=1
        if t[2]<=0:
            stage_content[t[1]]=t[0]; full[t[1]].arrive(); inflight.remove(t)
    order=list(range(5)); random.shuffle(order)
    for k in order:
        if alive[k]:
            try: next(procs[k])
            except StopIteration: alive[k]=False
else:
    print("ok", steps)
