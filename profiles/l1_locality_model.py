import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np
from mpas_regent_b200 import icosa, partition
n = int(sys.argv[1]) if len(sys.argv) > 1 else 163842
t0 = time.time()
mesh = icosa.make_icosahedral_mesh(n)
v = mesh.v
nC, nE = mesh.nCells, mesh.nEdges
ck = partition.hilbert_keys(v["xCell"], v["yCell"], v["zCell"])
cNew = np.empty(nC, np.int64); cNew[np.argsort(ck, kind="stable")] = np.arange(nC)
coe = v["cellsOnEdge"] - 1
a, b = cNew[coe[:, 0]], cNew[coe[:, 1]]
key_pair = (np.minimum(a, b) << 32) | np.maximum(a, b)
ek = partition.hilbert_keys(v["xEdge"], v["yEdge"], v["zEdge"])
orders = {"(min,max) cell pair [shipped]": np.argsort(key_pair, kind="stable"), "Hilbert key of the edge midpoint": np.argsort(ek, kind="stable")}
eoe = v["edgesOnEdge"] - 1
ne = v["nEdgesOnEdge"]
CPB = 8
print(f"mesh x1.{n}: {nE} edges; built in {time.time()-t0:.1f}s")
for name, order in orders.items():
    eNew = np.empty(nE, np.int64); eNew[order] = np.arange(nE)
    # neighbours of each edge in new numbering, rows in new order
    nb = np.where(np.arange(eoe.shape[1])[None, :] < ne[:, None], eNew[np.clip(eoe, 0, nE - 1)], -1)[order]
    d = np.abs(nb - np.arange(nE)[:, None]); d = d[nb >= 0]
    # distinct neighbour columns per block of CPB consecutive edges (one thread block), and per 4 blocks (L1-resident set)
    def distinct(rows):
        m = (nE // rows) * rows
        x = nb[:m].reshape(nE // rows, -1)
        x = np.sort(x, axis=1)
        return ((x[:, 1:] != x[:, :-1]) & (x[:, 1:] >= 0)).sum(1).mean() + 1
    print(f"{name:36s} median |eoe-e| {np.median(d):8.0f}  p90 {np.percentile(d,90):9.0f}  within 64: {np.mean(d<=64)*100:5.1f}%  within 1024: {np.mean(d<=1024)*100:5.1f}%"
          f"  distinct columns per 8-edge block {distinct(8):5.1f} (of {8*10}), per 64 edges {distinct(64):6.1f} (of {64*10})")

# LRU simulation: one block walks CONSECUTIVE 8-edge tiles (chunked persistent mapping) with room for `cap` neighbour columns
order = orders["(min,max) cell pair [shipped]"]
eNew = np.empty(nE, np.int64); eNew[order] = np.arange(nE)
nb = np.where(np.arange(eoe.shape[1])[None, :] < ne[:, None], eNew[np.clip(eoe, 0, nE - 1)], -1)[order]
from collections import OrderedDict
start = nE // 3
for cap in (40, 80, 160, 320):
    lru = OrderedDict(); hits = tot = 0
    for e in range(start, start + 40000):
        for x in nb[e]:
            if x < 0: continue
            tot += 1
            if x in lru: hits += 1; lru.move_to_end(x)
            else:
                lru[x] = 1
                if len(lru) > cap: lru.popitem(last=False)
    print(f"chunked walk, LRU capacity {cap:4d} columns: hit rate {hits/tot*100:5.1f}%")
# today's mapping: every tile starts cold (consecutive tiles run on different SMs)
hits = tot = 0
for t in range(start // 8, start // 8 + 5000):
    seen = set()
    for e in range(t * 8, t * 8 + 8):
        for x in nb[e]:
            if x < 0: continue
            tot += 1
            if x in seen: hits += 1
            else: seen.add(x)
print(f"one tile per block, cold start: hit rate {hits/tot*100:5.1f}%")
