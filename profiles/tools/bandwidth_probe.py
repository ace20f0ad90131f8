import gds_b200
ctx = gds_b200.Context(0)
n = 50_000_000
for r, w in [(1, 1), (2, 1), (3, 2), (5, 2), (4, 2), (8, 2), (4, 0), (1, 0), (0, 2)]:
    print(f'reads {r} writes {w}: {ctx.bandwidth_probe(r, w, n, 10):.0f} GB/s')
