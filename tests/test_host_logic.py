"""CPU: host-side logic that mirrors reference behaviour and needs no device — the LR schedule of
train.py:333-340/:375 against torch's own MultiStepLR, AugmentNoise's generator side effects and
statistics (train.py:64-94, training_script.md:4-10), the data-parallel slicing of noise / selectors."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("n_epoch", [10, 37, 100, 250])
def test_multistep_lr_equals_torch_multisteplr(n_epoch):
    """The reference builds MultiStepLR(milestones=[int(20r)-1, ...], gamma) and calls scheduler.step() at the
    end of every epoch (train.py:333-340, :375): the LR used DURING 1-based epoch e must match."""
    from image_denoising_b200.optim import multistep_lr, multistep_milestones
    from oracle import n2n_oracle as O
    w = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([w], lr=3e-4)
    ratio = n_epoch / 100
    ms = [int(20 * ratio) - 1, int(40 * ratio) - 1, int(60 * ratio) - 1, int(80 * ratio) - 1]
    assert multistep_milestones(n_epoch) == ms
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=ms, gamma=0.5)
    for epoch in range(1, n_epoch + 1):
        ref = opt.param_groups[0]["lr"]
        assert multistep_lr(3e-4, epoch, n_epoch, 0.5) == pytest.approx(ref, rel=1e-12), (n_epoch, epoch)
        assert O.multistep_lr(3e-4, epoch, n_epoch, 0.5) == pytest.approx(ref, rel=1e-12)
        opt.step()
        sched.step()


def test_multistep_lr_matches_reference_golden(golden):
    from image_denoising_b200.optim import multistep_lr
    lrs = golden("adam")["lrs_nepoch10"]
    assert np.allclose([multistep_lr(3e-4, e, 10, 0.5) for e in range(1, 11)], lrs, rtol=1e-12)


def test_train_entry_uses_the_shared_schedule():
    """entry/train.py once carried a private copy that decayed one epoch late."""
    import entry.train as T
    from image_denoising_b200 import optim
    assert T.multistep_lr is optim.multistep_lr
    # n_epoch=100, gamma=0.5: epochs 1..19 at base LR, 20..39 halved, ... (torch MultiStepLR, milestones 19/39/59/79)
    assert optim.multistep_lr(3e-4, 19, 100, 0.5) == 3e-4
    assert optim.multistep_lr(3e-4, 20, 100, 0.5) == 1.5e-4
    assert optim.multistep_lr(3e-4, 80, 100, 0.5) == pytest.approx(3e-4 / 16)


def test_augment_noise_gauss_counter_and_statistics():
    """train.py:84-94: every add_train_noise call takes a FRESH generator seeded with ++operation_seed_counter;
    gauss25 = N(0, (25/255)^2) added without clamping."""
    from image_denoising_b200 import n2n
    from oracle import n2n_oracle as O
    x = torch.rand(4, 1, 128, 128, generator=torch.Generator().manual_seed(1))
    n2n.operation_seed_counter = 10
    aug = n2n.AugmentNoise("gauss25")
    assert aug.style == "gauss_fix" and aug.params == [25 / 255.0]
    y1 = aug.add_train_noise(x)
    assert n2n.operation_seed_counter == 11
    y2 = aug.add_train_noise(x)
    assert n2n.operation_seed_counter == 12
    assert not torch.equal(y1, y2)
    assert torch.equal(y1, O.add_train_noise_gauss(x, 25.0, seed=11))          # same stream as the restatement
    n2n.operation_seed_counter = 10
    assert torch.equal(aug.add_train_noise(x), y1)                              # seed = counter value only
    noise = (y1 - x).double()
    assert abs(float(noise.std()) - 25 / 255.0) < 1e-3 and abs(float(noise.mean())) < 1e-3
    assert float(y1.min()) < 0.0 and float(y1.max()) > 1.0                      # no clamp (train.py:94)
    r = n2n.AugmentNoise("gauss5_50")
    assert r.style == "gauss_range" and r.params == [5 / 255.0, 50 / 255.0]
    yr = r.add_train_noise(x)
    per_sample = (yr - x).flatten(1).std(dim=1)
    assert ((per_sample > 4 / 255.0) & (per_sample < 51 / 255.0)).all()
    with pytest.raises(ValueError):
        n2n.AugmentNoise("speckle")


def test_data_parallel_noise_and_selector_are_slices_of_the_global_draw():
    """W ranks must see the noise / masks one process would have drawn for the concatenated batch."""
    from image_denoising_b200 import dp, n2n
    xg = torch.rand(8, 1, 32, 32, generator=torch.Generator().manual_seed(2))
    n2n.operation_seed_counter = 0
    ref = n2n.AugmentNoise("gauss25").add_train_noise(xg)
    parts = []
    for rank in range(4):
        n2n.operation_seed_counter = 0
        lo, hi = dp.shard_range(rank, 4, 8)
        parts.append(n2n.AugmentNoise("gauss25", rank=rank, world=4).add_train_noise(xg[lo:hi]))
    assert torch.equal(torch.cat(parts), ref)
    n2n.operation_seed_counter = 5
    rd_ref = n2n.draw_rd_idx(xg)
    got = []
    for rank in range(4):
        n2n.operation_seed_counter = 5
        lo, hi = dp.shard_range(rank, 4, 8)
        got.append(dp.shard_selector(n2n.draw_rd_idx(xg[lo:hi], batch=8), rank, 4, 8))
    assert torch.equal(torch.cat(got), rd_ref)


def test_adapter_refuses_gradient_through_base_out():
    """adapter.py:59-67 allows freeze_base=False / use_no_grad_for_base=False; the B200 path implements the frozen
    finetune only and must fail loudly rather than silently detach the base network."""
    from image_denoising_b200 import OutputAdapter
    ad = OutputAdapter(in_channels=1, hidden_channels=16)
    x = torch.zeros(1, 1, 8, 8)
    b = torch.zeros(1, 1, 8, 8, requires_grad=True)
    with pytest.raises(Exception) as e:
        ad(x, b)
    assert "not implemented" in str(e.value).lower() or "cuda" in str(e.value).lower()


def test_modules_copy_and_pickle_without_native_handles():
    """Plan handles / workspaces are per-process native objects: a deepcopy (EMA copies, DataParallel replicas) or a pickle
    of a module carries the parameters and starts with empty plan caches."""
    import copy
    import pickle
    from image_denoising_b200.adapter import OutputAdapter
    from image_denoising_b200.arch_unet import RESNET, UNet
    from image_denoising_b200.improved import ImprovedUNet
    for m in (UNet(1, 1, 4), RESNET(1, 1, 4), ImprovedUNet(1, 1, 16), OutputAdapter(1)):
        for c in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
            assert list(c.state_dict()) == list(m.state_dict())
            for a, b in zip(m.state_dict().values(), c.state_dict().values()):
                assert torch.equal(a, b)
            assert not getattr(c, "_plans", {}) and not getattr(c, "_free_ws", {})


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import or execute it.
    The package, the entry points and the native sources must not mention it in code (comments / docstrings may cite it)."""
    import ast
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for sub in ("image_denoising_b200", "entry"):
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                path = os.path.join(dirpath, f)
                if f.endswith(".py"):
                    tree = ast.parse(open(path).read())
                    for node in ast.walk(tree):
                        names = []
                        if isinstance(node, ast.Import):
                            names = [a.name for a in node.names]
                        elif isinstance(node, ast.ImportFrom):
                            names = [node.module or ""]
                        if any(n == "oracle" or n.startswith("oracle.") for n in names):
                            offenders.append(path)
                elif f.endswith((".cu", ".cuh", ".h")):
                    if any("oracle/" in line and "#include" in line for line in open(path)):
                        offenders.append(path)
    assert not offenders, offenders
    # bench.py may, in its BASELINE legs only: the CPU reference / port, and the oracle's functional graph run by stock PyTorch on
    # the GPU (torch_gpu_baseline and the stock-PyTorch comparison inside improved_unet_leg) — never in the measured product step
    src = open(os.path.join(root, "bench.py")).read()
    tree = ast.parse(src)
    users = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                users.add(fn.name)
    assert users <= {"cpu_reference_steps", "torch_gpu_baseline", "improved_unet_leg"}, users
    assert "run_b200" not in users
