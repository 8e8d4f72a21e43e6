"""Device-side initial-condition generators: the benchmark ICs of
SURVEY.md section 8d drawn directly into HBM with torch's CUDA generator, for use with
``ApproxState.from_device`` (no host arrays, so N ~ 1e9 per GPU is practical).

The draws differ from the numpy ``RandomState`` streams of the reference notebooks; use the host
generators in tests that compare against the reference.
"""


def sech2_disk(n, seed=2, zh=1., sigma=1., device='cuda'):
    """Isothermal sech^2 sheet (reference examples/WendyScaling.ipynb:57-65): returns (x, v, m0)."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    u = torch.rand(n, generator=g, dtype=torch.float64, device=device).clamp_(1e-300, 1. - 1e-16)
    x = torch.atanh(2. * u - 1.) * (2. * zh)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=device) * sigma
    v -= v.mean()
    return x, v, 1. / n


def cold_slab(n, seed=3, width=1., sigma=0.05, device='cuda'):
    """Cold uniform slab that collapses and phase-mixes (BASELINE config 2): returns (x, v, m0)."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    x = (torch.rand(n, generator=g, dtype=torch.float64, device=device) - 0.5) * width
    v = torch.randn(n, generator=g, dtype=torch.float64, device=device) * sigma
    v -= v.mean()
    return x, v, 1. / n


def exponential_disk(n, seed=4, zh=1., sigma=1., device='cuda'):
    """Double-exponential sheet (reference examples/AdiabaticVsNonAdiabatic.ipynb): returns (x, v, m0)."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    u = torch.rand(n, generator=g, dtype=torch.float64, device=device).clamp_(1e-300, 1.)
    sgn = torch.where(torch.rand(n, generator=g, dtype=torch.float64, device=device) < 0.5, -1., 1.)
    x = -zh * torch.log(u) * sgn
    v = torch.randn(n, generator=g, dtype=torch.float64, device=device) * sigma
    v -= v.mean()
    return x, v, 1. / n
