def natsorted(x, *a, **k): return sorted(x)
