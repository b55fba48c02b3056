"""CPU: host-side logic of the product package that needs no device."""
import torch


def test_packed_weight_cache_follows_every_kind_of_update():
    """Fused optimizers update parameters WITHOUT bumping Tensor._version; the packed bf16 copies must still be rebuilt
    (a stale forward weight silently stops training).  Also: in-place edits, storage replacement, per-object cache."""
    from cooperative_training_and_latent_space_data_augmentation_b200 import fastpath
    p = torch.nn.Parameter(torch.randn(4, 4))
    calls = []

    def pack(w):
        calls.append(1)
        return w.detach().clone()

    a = fastpath._packed(p, pack)
    assert fastpath._packed(p, pack) is a and len(calls) == 1
    assert fastpath._packed(p, pack, tag='dgrad') is not a and len(calls) == 2          # tags are separate entries
    opt = torch.optim.Adam([p], lr=1e-2, fused=True)
    p.grad = torch.ones_like(p)
    opt.step()
    b = fastpath._packed(p, pack)
    assert len(calls) == 3 and not torch.equal(a, b) and torch.equal(b, p.detach())
    with torch.no_grad():
        p.mul_(2.0)
    assert torch.equal(fastpath._packed(p, pack), p.detach()) and len(calls) == 4
    p.data = torch.zeros(4, 4)
    assert torch.equal(fastpath._packed(p, pack), torch.zeros(4, 4)) and len(calls) == 5
    q = torch.nn.Parameter(torch.ones(4, 4))                                               # a different object never hits
    assert torch.equal(fastpath._packed(q, pack), torch.ones(4, 4)) and len(calls) == 6
    fastpath.weights_changed()
    fastpath._packed(q, pack)
    assert len(calls) == 7


def test_batched_pack_registry_refreshes_every_weight_in_one_launch(monkeypatch):
    """fastpath._PackRegistry (the conv weights' forward / input-gradient packings): first request packs alone and
    registers a persistent buffer; after an optimizer step ONE batched launch refreshes every registered packing (job
    table with one row per entry); a moved storage re-registers; the table must exist before a CUDA-graph capture."""
    from cooperative_training_and_latent_space_data_augmentation_b200 import fastpath, ops
    single, batched, capturing = [], [], [False]

    def fake_pack_kernel(param, transposed, tap_major=None):
        single.append((id(param), transposed))
        return param.detach().clone().reshape(-1)            # persistent "packed" buffer

    def fake_pack_job(weight, transposed, out, tap_major=None):
        return [weight.data_ptr(), out.data_ptr(), weight.shape[0], weight.shape[1], 9, 16, int(transposed), 0]

    def fake_batched(table, n_jobs, max_elements):
        assert table.dtype == torch.int64 and tuple(table.shape) == (n_jobs, 8)
        batched.append((n_jobs, max_elements, table.clone()))

    monkeypatch.setattr(ops, "_pack_kernel", fake_pack_kernel)
    monkeypatch.setattr(ops, "pack_job", fake_pack_job)
    monkeypatch.setattr(ops, "pack_conv_weights_batched", fake_batched)
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: capturing[0])
    reg = fastpath._PackRegistry()
    w1 = torch.nn.Parameter(torch.randn(16, 16, 3, 3))
    w2 = torch.nn.Parameter(torch.randn(32, 16, 3, 3))
    a1 = reg.get(w1, False, 'fwd')
    a1d = reg.get(w1, True, 'dgrad')
    a2 = reg.get(w2, False, 'fwd')
    assert len(single) == 3 and not batched and reg.dirty
    assert reg.get(w1, False, 'fwd') is a1 and len(single) == 3                 # cached within the same weights epoch
    fastpath.weights_changed()                                                   # = an optimizer stepped
    assert reg.get(w2, False, 'fwd') is a2                                      # same persistent buffer ...
    assert len(batched) == 1 and batched[0][0] == 3 and batched[0][1] == w2.numel()   # ... refreshed with ALL others
    rows = batched[0][2]
    assert sorted(rows[:, 0].tolist()) == sorted([w1.data_ptr(), w1.data_ptr(), w2.data_ptr()])
    assert sorted(rows[:, 1].tolist()) == sorted([a1.data_ptr(), a1d.data_ptr(), a2.data_ptr()])
    assert reg.get(w1, False, 'fwd') is a1 and reg.get(w1, True, 'dgrad') is a1d and len(batched) == 1
    assert len(single) == 3 and not reg.dirty
    with torch.no_grad():
        w1.mul_(2.0)                                                             # in-place edit (version bump) -> refresh
    reg.get(w1, False, 'fwd')
    assert len(batched) == 2
    w2.data = torch.zeros(32, 16, 3, 3)                                          # storage moved -> packed alone, re-registered
    b2 = reg.get(w2, False, 'fwd')
    assert len(single) == 4 and b2 is not a2 and reg.dirty
    # a capture must find the table built: a stale table raises instead of copying host -> device inside the capture
    fastpath.weights_changed()
    capturing[0] = True
    try:
        reg.get(w1, False, 'fwd')
        raised = False
    except RuntimeError:
        raised = True
    assert raised
    capturing[0] = False
    reg._rebuild_table()
    capturing[0] = True
    reg.get(w1, False, 'fwd')                                                    # table ready: the launch is capturable
    assert len(batched) == 3 and batched[2][0] == 3
    w3 = torch.nn.Parameter(torch.randn(16, 16, 1, 1))
    try:
        reg.get(w3, False, 'fwd')                                                # first sight of a weight inside a capture
        raised = False
    except RuntimeError:
        raised = True
    assert raised


def test_pack_registry_table_never_names_dead_or_moved_weights(monkeypatch):
    """The job table holds RAW pointers: (1) it pins the storages it names (a model that dies while the table -- or a
    retired table a captured graph still reads -- is alive cannot leave dangling rows), (2) a rebuild drops the entries of
    dead parameters together with their packed buffers, (3) a refresh outside a capture notices a registered weight that
    died or moved since the table was built and rebuilds first."""
    import gc
    from cooperative_training_and_latent_space_data_augmentation_b200 import fastpath, ops
    batched = []
    monkeypatch.setattr(ops, "_pack_kernel", lambda param, transposed, tap_major=None: param.detach().clone().reshape(-1))
    monkeypatch.setattr(ops, "pack_job", lambda weight, transposed, out, tap_major=None:
                        [weight.data_ptr(), out.data_ptr(), weight.shape[0], weight.shape[1], 9, 16, int(transposed), 0])
    monkeypatch.setattr(ops, "pack_conv_weights_batched", lambda table, n, m: batched.append(table.clone()))
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)
    reg = fastpath._PackRegistry()
    w1 = torch.nn.Parameter(torch.randn(16, 16, 3, 3))
    w2 = torch.nn.Parameter(torch.randn(16, 16, 3, 3))
    reg.get(w1, False, 'fwd')
    reg.get(w2, False, 'fwd')
    reg._rebuild_table()
    assert reg._table_is_current() and len(reg.sources) == 2
    p1 = w1.data_ptr()
    assert any(src.data_ptr() == p1 for src in reg.sources)                      # (1) the storage is pinned by the table
    del w1
    gc.collect()
    assert not reg._table_is_current()                                           # (3) a dead weight makes the table stale
    fastpath.weights_changed()
    reg.get(w2, False, 'fwd')                                                    # refresh: rebuilds first, then ONE launch
    assert len(batched) == 1 and batched[0].shape[0] == 1 and int(batched[0][0, 0]) == w2.data_ptr()
    assert len(reg.entries) == 1 and reg._table_is_current()                     # (2) the dead entry is gone
    assert reg.retired and any(src.data_ptr() == p1 for src in reg.retired[-1][1])   # the retired table keeps ITS sources
    w2.data = torch.randn(16, 16, 3, 3)                                          # storage moved behind the registry's back
    assert not reg._table_is_current()
    fastpath.prepare_packing.__globals__['_REGISTRY'], saved = reg, fastpath.prepare_packing.__globals__['_REGISTRY']
    try:
        fastpath.prepare_packing()                                               # what a trainer calls before a capture
        assert reg._table_is_current() and int(reg.table[0, 0]) == w2.data_ptr()
    finally:
        fastpath.prepare_packing.__globals__['_REGISTRY'] = saved


def test_grads_direct_mode_accumulates_in_place_and_returns_none():
    """trainpath._Grads: inside accumulate_into_grads() parameters that own a .grad receive their gradients in place
    (kernel-accumulated buffers are the .grad itself, small vectors go through one foreach add, a second contribution
    to the same parameter is added immediately) and autograd gets None; outside, or when a .grad is missing, gradients
    are returned and summed per parameter."""
    from cooperative_training_and_latent_space_data_augmentation_b200 import trainpath
    w = torch.nn.Parameter(torch.zeros(2, 3, 1, 1))
    b = torch.nn.Parameter(torch.zeros(2))
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    params, needs = (w, b, frozen), (True, True, False)
    # ordinary mode: returned, summed, frozen parameter ignored
    g = trainpath._Grads(params, needs)
    assert not g.direct
    buf = g.buffer(w)
    assert buf.shape == w.shape and float(buf.abs().sum()) == 0.0
    buf += 1.0
    g.add(w, buf)
    g.add(b, torch.ones(2)); g.add(b, 2 * torch.ones(2)); g.add(frozen, torch.ones(2)); g.add(b, None)
    out = g.results(params)
    assert torch.equal(out[0], torch.ones_like(w)) and torch.equal(out[1], 3 * torch.ones(2)) and out[2] is None
    # direct mode needs the context AND existing .grad tensors
    with trainpath.accumulate_into_grads():
        assert not trainpath._Grads(params, needs).direct            # no .grad yet
    w.grad, b.grad = torch.full_like(w, 10.0), torch.full_like(b, 10.0)
    assert not trainpath._Grads(params, needs).direct                # .grad exists but the context is off
    with trainpath.accumulate_into_grads():
        g = trainpath._Grads(params, needs)
        assert g.direct and g.arena is None
        buf = g.buffer(w)
        assert buf.data_ptr() == w.grad.data_ptr()                   # kernels accumulate into .grad itself
        buf += 1.0
        g.add(w, buf)                                                # ... and that is not added a second time
        g.add(b, torch.ones(2)); g.add(b, 2 * torch.ones(2)); g.add(frozen, torch.ones(2))
        out = g.results(params)
    assert out == (None, None, None)
    assert torch.equal(w.grad, torch.full_like(w, 11.0)) and torch.equal(b.grad, torch.full_like(b, 13.0))
    assert not trainpath._DIRECT["on"]


def test_identities_the_kernel_path_relies_on():
    """(1) a 1x1 convolution commutes with nearest-neighbour up-sampling (trainpath.residual_fwd takes the shortcut of
    res_up_family, encoder_decoder.py:294-296 / :334-337, at the low resolution); (2) the (hi | lo | hi) x (w_hi | w_hi |
    w_lo) channel-group layout of the tensor-core stem reproduces the fp32 product to ~2^-16 although every operand the
    tensor core sees is bf16 (ops.pad_stem_weight / ctl_stem_input_c8)."""
    import torch.nn.functional as F
    from cooperative_training_and_latent_space_data_augmentation_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 5, 7, generator=g)
    w = torch.randn(4, 8, 1, 1, generator=g)
    b = torch.randn(4, generator=g)
    up = lambda t: F.interpolate(t, scale_factor=2, mode='nearest')
    assert torch.equal(F.conv2d(up(x), w, b), up(F.conv2d(x, w, b)))

    cin = 4
    xs = torch.rand(2, cin, 9, 11, generator=g)
    ws = 0.3 * torch.randn(16, cin, 3, 3, generator=g)
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    hi = bf(xs)
    x16 = torch.zeros(2, 16, 9, 11)
    x16[:, :cin], x16[:, cin:2 * cin], x16[:, 2 * cin:3 * cin] = hi, bf(xs - hi), hi       # what ctl_stem_input_c8 writes
    w16 = bf(ops.pad_stem_weight(ws))                                                     # the packing rounds to bf16
    want = F.conv2d(xs.double(), ws.double(), padding=1)
    got = F.conv2d(x16.double(), w16.double(), padding=1)
    plain = F.conv2d(hi.double(), bf(ws).double(), padding=1)                             # single bf16 operands
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) < 3e-5 * scale
    assert float((plain - want).abs().max()) > 20 * float((got - want).abs().max())       # what the split buys
    # the weight gradient's two channel groups add up to sum dy * x
    dy = torch.randn(2, 16, 9, 11, generator=g)
    xi = x16.clone().requires_grad_(False)
    w_var = torch.zeros(16, 16, 3, 3, requires_grad=True)
    F.conv2d(xi, w_var, padding=1).backward(dy)
    w_ref = ws.clone().requires_grad_(True)
    F.conv2d(xs, w_ref, padding=1).backward(dy)
    err = (ops.stem_weight_grad(w_var.grad, cin) - w_ref.grad).abs().max()
    assert float(err) < 3e-5 * float(w_ref.grad.abs().max())
