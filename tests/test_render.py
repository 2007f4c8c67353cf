"""Permittivity renderer: the golden values of
/root/reference/tests/test_layers.py, checked on the NumPy oracle (CPU) and on the CUDA renderer
(``pjz_b200._epsilon.render``), plus CUDA-vs-oracle on random layer stacks."""

import numpy as np
import pytest

from oracle import render_numpy

Z = lambda zz: (np.arange(zz)[:, None] + np.array([[-0.5, 0]]), np.arange(zz)[:, None] + np.array([[0.5, 1]]))


def _impls():
  return [pytest.param("oracle", id="oracle"),
          pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def _render(impl, layers, pos, gs, ge, m, simple=False):
  if impl == "oracle":
    return render_numpy.render(layers, pos, gs, ge, m, simple)
  from pjz_b200._epsilon import render
  return render(np.asarray(layers, np.float32), np.asarray(pos, np.float32), gs, ge, m, simple).cpu().numpy()


SIMPLE = [
    ((5, 0, -0.5), (0, 1, 0, 0), 1.0), ((5, 0, -0.5), (0, 2, 0, 0), 1.5), ((5, 0, -0.5), (0, 3, 0, 0), 3.0),
    ((4, 0, -0.5), (1, 1, 0, 0), 1.0), ((4, 0, -0.5), (1, 2, 0, 0), 2.0), ((4, 0, -0.5), (1, 3, 0, 0), 3.0),
    ((0, 5, -0.5), (1, 0, 1, 0), 1.0), ((0, 5, -0.5), (1, 0, 2, 0), 1.5), ((0, 5, -0.5), (1, 0, 3, 0), 3.0),
    ((0, 4, -0.5), (2, 0, 1, 0), 1.0), ((0, 4, -0.5), (2, 0, 2, 0), 2.0), ((0, 4, -0.5), (2, 0, 3, 0), 3.0),
    ((0, 0, +2.5), (2, 0, 0, 1), 1.0), ((0, 0, +2.5), (2, 0, 0, 2), 1.5), ((0, 0, +2.5), (2, 0, 0, 3), 3.0),
    ((0, 0, +2.0), (1, 0, 0, 1), 1.0), ((0, 0, +2.0), (1, 0, 0, 2), 2.0), ((0, 0, +2.0), (1, 0, 0, 3), 3.0),
]


@pytest.mark.parametrize("impl", _impls())
def test_simple(impl):
  # /root/reference/tests/test_layers.py:8-40
  xx, yy, zz = 4, 5, 6
  for thresh, index, expected in SIMPLE:
    layer = np.ones((2, 2 * xx, 2 * yy))
    layer[1, thresh[0]:, thresh[1]:] = 3
    out = _render(impl, layer, [thresh[2]], *Z(zz), 1)
    assert out[index] == pytest.approx(expected, rel=1e-6), (thresh, index)


X_THRESHOLD = [
    (1, 3, 0, 0, 1.0), (1, 3, 0, 1, 1.5), (1, 3, 0, 2, 3.0), (2, 6, 0, 0, 1.0), (2, 6, 0, 1, 1.5),
    (2, 6, 0, 2, 3.0), (3, 9, 0, 0, 1.0), (3, 9, 0, 1, 1.5), (3, 9, 0, 2, 3.0),
    (3, 6, 0, 1, 3.0), (3, 7, 0, 1, 9 / 4), (3, 8, 0, 1, 9 / 5), (3, 9, 0, 1, 1.5), (3, 10, 0, 1, 9 / 7),
    (3, 11, 0, 1, 9 / 8), (3, 12, 0, 1, 1.0),
    (1, 4, 1, 1, 1.0), (1, 4, 1, 2, 2.0), (1, 4, 1, 3, 3.0), (2, 8, 1, 1, 1.0), (2, 8, 1, 2, 2.0),
    (2, 8, 1, 3, 3.0), (3, 12, 1, 1, 1.0), (3, 12, 1, 2, 2.0), (3, 12, 1, 3, 3.0),
    (3, 9, 1, 2, 3.0), (3, 10, 1, 2, 8 / 3), (3, 11, 1, 2, 7 / 3), (3, 12, 1, 2, 2.0), (3, 13, 1, 2, 5 / 3),
    (3, 14, 1, 2, 4 / 3), (3, 15, 1, 2, 1.0),
]


@pytest.mark.parametrize("impl", _impls())
def test_x_threshold(impl):
  # /root/reference/tests/test_layers.py:43-91
  xx, yy, zz = 4, 3, 2
  for m, thresh, component, index, expected in X_THRESHOLD:
    layer = np.ones((1, 2 * m * xx, 2 * m * yy))
    layer[0, thresh:, :] = 3
    out = _render(impl, layer, [], *Z(zz), m)
    assert out[component, index, 0, 0] == pytest.approx(expected, rel=1e-6), (m, thresh, component, index)


Z_THRESHOLD = [(0.00, 2, 0, 3.0), (0.25, 2, 0, 2.0), (0.50, 2, 0, 1.5), (0.75, 2, 0, 1.2), (1.00, 2, 0, 1.0),
               (-0.50, 0, 0, 3.0), (-0.25, 0, 0, 2.5), (+0.00, 0, 0, 2.0), (+0.25, 0, 0, 1.5),
               (+0.50, 0, 0, 1.0)]


@pytest.mark.parametrize("impl", _impls())
def test_z_threshold(impl):
  # /root/reference/tests/test_layers.py:94-119
  for thresh, component, index, expected in Z_THRESHOLD:
    layer = np.ones((2, 2, 2))
    layer[1] = 3
    out = _render(impl, layer, [thresh], *Z(1), 1)
    assert out[component, 0, 0, index] == pytest.approx(expected, rel=1e-6), (thresh, component)


A, B3 = 1 / (0.5 * (5 / 6) + 0.5 * (2 / 3)), 1 / ((1 / 3) * (11 / 12) + (2 / 3) * (4 / 5))
CORNER = [
    ((0, 0, -0.5), (0, 0, 1, 0), 3.0), ((1, 0, -0.5), (0, 0, 1, 0), 1.5), ((2, 0, -0.5), (0, 0, 1, 0), 1.0),
    ((0, 2, -0.5), (0, 0, 1, 0), 2.0), ((0, 3, -0.5), (0, 0, 1, 0), 1.0), ((0, 0, +0.0), (0, 0, 1, 0), 2.0),
    ((0, 0, +0.5), (0, 0, 1, 0), 1.0), ((1, 2, -0.5), (0, 0, 1, 0), A), ((0, 2, +0.0), (0, 0, 1, 0), 1.5),
    ((1, 0, +0.0), (0, 0, 1, 0), A), ((1, 2, +0.0), (0, 0, 1, 0), B3),
    ((0, 0, -0.5), (1, 1, 0, 0), 3.0), ((2, 0, -0.5), (1, 1, 0, 0), 2.0), ((3, 0, -0.5), (1, 1, 0, 0), 1.0),
    ((0, 1, -0.5), (1, 1, 0, 0), 1.5), ((0, 2, -0.5), (1, 1, 0, 0), 1.0), ((0, 0, +0.0), (1, 1, 0, 0), 2.0),
    ((0, 0, +0.5), (1, 1, 0, 0), 1.0), ((2, 1, -0.5), (1, 1, 0, 0), A), ((0, 1, +0.0), (1, 1, 0, 0), A),
    ((2, 0, +0.0), (1, 1, 0, 0), 1.5), ((2, 1, +0.0), (1, 1, 0, 0), B3),
    ((0, 0, 0.0), (2, 1, 1, 0), 3.0), ((2, 0, 0.0), (2, 1, 1, 0), 2.0), ((3, 0, 0.0), (2, 1, 1, 0), 1.0),
    ((0, 2, 0.0), (2, 1, 1, 0), 2.0), ((0, 3, 0.0), (2, 1, 1, 0), 1.0), ((0, 0, 0.5), (2, 1, 1, 0), 1.5),
    ((0, 0, 1.0), (2, 1, 1, 0), 1.0), ((2, 2, 0.0), (2, 1, 1, 0), 1.5), ((0, 2, 0.5), (2, 1, 1, 0), A),
    ((2, 0, 0.5), (2, 1, 1, 0), A), ((2, 2, 0.5), (2, 1, 1, 0), B3),
]


@pytest.mark.parametrize("impl", _impls())
def test_corner(impl):
  # /root/reference/tests/test_layers.py:122-176
  for thresh, index, expected in CORNER:
    layer = np.ones((2, 4, 4))
    layer[1, thresh[0]:, thresh[1]:] = 3
    out = _render(impl, layer, [thresh[2]], *Z(2), 1)
    assert out[index] == pytest.approx(expected, rel=1e-6), (thresh, index)


@pytest.mark.gpu
@pytest.mark.parametrize("ll,xx,yy,zz,m,simple", [(1, 5, 4, 3, 1, False), (3, 9, 7, 12, 2, False),
                                                 (4, 16, 12, 20, 4, False), (3, 6, 5, 8, 3, True)])
def test_cuda_renderer_matches_the_oracle_on_random_stacks(ll, xx, yy, zz, m, simple, built):
  rng = np.random.default_rng(ll * 100 + m)
  layers = rng.uniform(1.0, 12.25, (ll, 2 * m * xx, 2 * m * yy)).astype(np.float32)
  pos = np.sort(rng.uniform(0.5, zz - 1.5, ll - 1)).astype(np.float32)
  gs = (np.arange(zz)[:, None] * 1.1 + np.array([[-0.5, 0]])).astype(np.float32)   # stretched grid
  ge = (np.arange(zz)[:, None] * 1.1 + np.array([[0.6, 1.1]])).astype(np.float32)
  want = render_numpy.render(layers, pos, gs, ge, m, simple)
  got = _render("cuda", layers, pos, gs, ge, m, simple)
  assert got.shape == (3, xx, yy, zz) and got.dtype == np.float32
  np.testing.assert_allclose(got, want, rtol=2e-6)


@pytest.mark.gpu
def test_epsilon_feeds_the_engine_without_leaving_the_gpu(built):
  """pjz.epsilon -> fdtdz_jax.fdtdz on device tensors: same snapshots as with a host epsilon."""
  import torch
  from pjz_b200 import fdtdz_jax
  from pjz_b200._epsilon import epsilon
  from tests.problems import random_problem
  rng = np.random.default_rng(5)
  layers = rng.uniform(1.0, 12.25, (3, 2 * 8, 2 * 7)).astype(np.float32)
  eps = epsilon(layers, np.array([3.2, 7.9], np.float32), 1, 11)
  assert eps.is_cuda and tuple(eps.shape) == (3, 8, 7, 11)
  kw = random_problem(domain=(12, 10, 16), sub=(8, 7, 11), offset=(2, 1, 3), tt=12, seed=3)
  kw["epsilon"] = eps
  a = fdtdz_jax.fdtdz(**kw).cpu().numpy()
  kw["epsilon"] = eps.cpu().numpy()
  np.testing.assert_array_equal(a, fdtdz_jax.fdtdz(**kw))


# ---- backward pass (the reference differentiates pjz.render with jax.grad,
# /root/reference/tests/test_layers.py:179-188) -------------------------------------------------------

def _random_stack(ll, xx, yy, zz, m, seed):
  rng = np.random.default_rng(seed)
  layers = rng.uniform(1.0, 12.25, (ll, 2 * m * xx, 2 * m * yy))
  pos = np.sort(rng.uniform(0.5, zz - 1.5, ll - 1))
  gs = np.arange(zz)[:, None] * 1.1 + np.array([[-0.5, 0]])          # stretched grid
  ge = np.arange(zz)[:, None] * 1.1 + np.array([[0.6, 1.1]])
  wts = rng.standard_normal((3, xx, yy, zz))                         # d loss / d epsilon
  return layers, pos, gs, ge, wts


@pytest.mark.parametrize("simple", [False, True])
def test_torch_oracle_equals_numpy_oracle_and_its_gradient_matches_differences(simple):
  """The differentiable float64 restatement (oracle/render_torch.py) agrees with the NumPy oracle
  -- which the golden values above pin -- and its autograd gradient with central differences of
  the NumPy oracle, for layers and layer_pos."""
  import torch
  from oracle import render_torch
  ll, xx, yy, zz, m = 3, 3, 2, 4, 2
  layers, pos, gs, ge, wts = _random_stack(ll, xx, yy, zz, m, 11)
  want = render_numpy.render(layers, pos, gs, ge, m, simple)
  lt = torch.tensor(layers, requires_grad=True)
  pt = torch.tensor(pos, requires_grad=True)
  got = render_torch.render(lt, pt, torch.tensor(gs), torch.tensor(ge), m, simple)
  np.testing.assert_allclose(got.detach().numpy(), want, rtol=1e-12)
  (got * torch.tensor(wts)).sum().backward()
  loss = lambda L, P: float((render_numpy.render(L, P, gs, ge, m, simple) * wts).sum())
  h = 1e-6
  rng = np.random.default_rng(0)
  for idx in [tuple(rng.integers(0, s) for s in layers.shape) for _ in range(25)] + \
             [(0, 0, 0), (1, 0, 3), (2, 5, 0), (0, 2 * m * xx - 1, 2 * m * yy - 1)]:
    d = np.zeros_like(layers); d[idx] = h
    fd = (loss(layers + d, pos) - loss(layers - d, pos)) / (2 * h)
    assert lt.grad[idx].item() == pytest.approx(fd, rel=2e-5, abs=1e-8), idx
  for k in range(ll - 1):
    d = np.zeros_like(pos); d[k] = h
    fd = (loss(layers, pos + d) - loss(layers, pos - d)) / (2 * h)
    assert pt.grad[k].item() == pytest.approx(fd, rel=2e-5, abs=1e-8), k


def test_torch_oracle_is_differentiable_like_the_reference():
  # /root/reference/tests/test_layers.py:179-188 on the oracle
  import torch
  from oracle import render_torch
  xx, yy, zz = 2, 2, 2
  lay = torch.ones((2, 2 * xx, 2 * yy), dtype=torch.float64, requires_grad=True)
  pos = torch.tensor([1.0], dtype=torch.float64, requires_grad=True)
  gs, ge = Z(zz)
  render_torch.render(lay, pos, torch.tensor(gs), torch.tensor(ge), 1).sum().backward()
  assert lay.grad.shape == (2, 2 * xx, 2 * yy) and pos.grad.shape == (1,)


@pytest.mark.gpu
def test_cuda_renderer_is_differentiable_like_the_reference(built):
  # /root/reference/tests/test_layers.py:179-188 on the CUDA renderer
  import torch
  from pjz_b200._epsilon import render
  xx, yy, zz = 2, 2, 2
  lay = torch.ones((2, 2 * xx, 2 * yy), device="cuda", requires_grad=True)
  pos = torch.tensor([1.0], device="cuda", requires_grad=True)
  gs, ge = Z(zz)
  render(lay, pos, gs, ge, 1).sum().backward()
  assert lay.grad.shape == (2, 2 * xx, 2 * yy) and pos.grad.shape == (1,)
  assert torch.isfinite(lay.grad).all() and torch.isfinite(pos.grad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("ll,xx,yy,zz,m,simple", [(1, 5, 4, 3, 1, False), (2, 2, 2, 2, 1, False),
                                                 (3, 9, 7, 12, 2, False), (4, 16, 12, 20, 4, False),
                                                 (3, 6, 5, 8, 3, True), (5, 33, 17, 96, 2, False)])
def test_cuda_renderer_backward_matches_the_torch_oracle(ll, xx, yy, zz, m, simple, built):
  """b200fdtd_render_backward (d/d layers, d/d layer_pos) against autograd through the float64
  torch restatement; tolerance 2e-5 of the gradient's largest entry (the forward's tile
  statistics are held in float32)."""
  import torch
  from oracle import render_torch
  from pjz_b200._epsilon import render
  layers, pos, gs, ge, wts = _random_stack(ll, xx, yy, zz, m, ll * 100 + m)
  lt = torch.tensor(layers, requires_grad=True)
  pt = torch.tensor(pos, requires_grad=True)
  (render_torch.render(lt, pt, torch.tensor(gs), torch.tensor(ge), m, simple) * torch.tensor(wts)).sum().backward()
  lc = torch.tensor(layers, dtype=torch.float32, device="cuda", requires_grad=True)
  pc = torch.tensor(pos, dtype=torch.float32, device="cuda", requires_grad=True)
  out = render(lc, pc, gs, ge, m, simple)
  (out * torch.tensor(wts, dtype=torch.float32, device="cuda")).sum().backward()
  gl, wl = lc.grad.cpu().double(), lt.grad
  assert gl.shape == wl.shape
  assert float((gl - wl).abs().max()) <= 2e-5 * float(wl.abs().max())
  if ll > 1:
    gp, wp = pc.grad.cpu().double(), pt.grad
    assert float((gp - wp).abs().max()) <= 2e-5 * float(wp.abs().max()) + 1e-7


@pytest.mark.gpu
def test_gradient_flows_from_the_engine_loss_to_the_layers(built):
  """layers -> epsilon (CUDA renderer) -> scatter (CUDA engine, fused adjoint backward) -> loss:
  one backward() reaches the layer image and the interface positions on the GPU."""
  import torch
  from pjz_b200 import SimParams, mode, scatter
  from pjz_b200._epsilon import epsilon
  omega = np.array([2 * np.pi / 37])
  xx, yy, zz = 40, 30, 20
  lay = np.full((3, 2 * xx, 2 * yy), 1.0, np.float32)
  lay[1, :, 2 * 9:2 * 21] = 12.25                        # Si strip in the middle layer
  lt = torch.tensor(lay, device="cuda", requires_grad=True)
  pos = torch.tensor([7.5, 11.5], device="cuda", requires_grad=True)
  eps = epsilon(lt, pos, 1, zz)
  assert eps.requires_grad and tuple(eps.shape) == (3, xx, yy, zz)
  e_np = eps.detach().cpu().numpy()
  b0, x0, _, _ = mode(e_np[:, 6:7], omega, 1)
  b1, x1, _, _ = mode(e_np[:, 33:34], omega, 1)
  p = SimParams(omega_range=(omega[0], omega[0]), tt=400, dt=0.5, absorption_padding=10,
                absorption_coeff=4e-4, pml_widths=(6, 6), use_reduced_precision=False, domain_zz=32)
  sv = scatter(eps, omega, [x0[..., 0], x1[..., 0]], [b0[:, 0], b1[:, 0]], [6, 33], [True, False], p,
               fuse_projection=True)
  loss = sum((s.abs() ** 2).sum() for row in sv for s in row)
  loss.backward()
  assert lt.grad is not None and lt.grad.shape == lt.shape and torch.isfinite(lt.grad).all()
  assert float(lt.grad.abs().max()) > 0
  assert pos.grad is not None and pos.grad.shape == (2,) and torch.isfinite(pos.grad).all()


def test_overlap_table_host_logic_matches_the_reference_formulas():
  """pjz_b200._epsilon._overlap_table (the host-side link between b200fdtd_render_backward's
  overlap-table gradient and d/d layer_pos) against /root/reference/src/pjz/_epsilon.py:55-66
  restated with NumPy, and its derivative against differences."""
  import torch
  from pjz_b200._epsilon import _overlap_table
  rng = np.random.default_rng(4)
  zz, ll = 9, 4
  pos = np.sort(rng.uniform(0.3, zz - 1.2, ll - 1))
  gs = np.arange(zz)[:, None] * 1.05 + np.array([[-0.5, 0]])
  ge = np.arange(zz)[:, None] * 1.05 + np.array([[0.55, 1.05]])

  def table(p):
    us, uzs = [], []
    for col in (0, 1):
      s, e = gs[:, col], ge[:, col]
      p0, p1 = [np.clip(x[:, None], s, e) for x in (np.concatenate([[-np.inf], p]),
                                                    np.concatenate([p, [np.inf]]))]
      u = (p1 - p0) / (e - s)
      us.append(u); uzs.append(u * ((p0 + p1) / 2 - (s + e) / 2))
    return np.stack(us), np.stack(uzs)

  pt = torch.tensor(pos, requires_grad=True)
  u, uz = _overlap_table(pt, torch.tensor(gs, dtype=torch.float32), torch.tensor(ge, dtype=torch.float32))
  wu, wuz = table(pos)
  assert tuple(u.shape) == (2, ll, zz)
  np.testing.assert_allclose(u.detach().numpy(), wu, atol=1e-6)
  np.testing.assert_allclose(uz.detach().numpy(), wuz, atol=1e-6)
  np.testing.assert_allclose(u.detach().numpy().sum(1), 1.0, atol=1e-12)     # layers tile every cell
  cu, cuz = rng.standard_normal(wu.shape), rng.standard_normal(wuz.shape)
  ((u * torch.tensor(cu)).sum() + (uz * torch.tensor(cuz)).sum()).backward()
  h = 1e-6
  for k in range(ll - 1):
    d = np.zeros_like(pos); d[k] = h
    (a0, b0), (a1, b1) = table(pos - d), table(pos + d)
    fd = (((a1 - a0) * cu).sum() + ((b1 - b0) * cuz).sum()) / (2 * h)
    assert pt.grad[k].item() == pytest.approx(fd, rel=1e-4, abs=1e-6)
