"""GPU: the loader / dump edges (csrc/io_edges.cu) against fixtures produced by the unmodified
reference — label-id encode (cityscapes.py:85-91), full_image_eval_preprocess (custom_transforms.py:322-347: normalise,
pad image with 0 and labels with 255), decode_segmap (dataloaders/utils.py:14-51), and the checkpoint loader
(eval.py:126-140).  Everything here is integer / byte work or the loader's exact float arithmetic: bit-exact."""
import numpy as np
import pytest
import torch

import util
import add_b200

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
IO = np.load(util.ROOT / "tests/golden/io_edges.npz")


def test_encode_segmap_all_ids():
    ids = torch.arange(256, dtype=torch.uint8).reshape(16, 16).to(DEV)
    assert np.array_equal(add_b200.encode_segmap(ids).cpu().numpy(), IO["encode/all_ids"])
    assert np.array_equal(add_b200.cityscapes_label_lut().reshape(16, 16), IO["encode/all_ids"])


@pytest.mark.parametrize("name", sorted(util.IO_CASES))
def test_eval_preprocess_and_decode(name):
    spec = util.IO_CASES[name]
    img, ids = util.make_io_case(name)
    img_d, ids_d = torch.from_numpy(img).to(DEV), torch.from_numpy(ids).to(DEV)
    enc = add_b200.encode_segmap(ids_d)
    assert np.array_equal(enc.cpu().numpy(), IO[f"{name}/encoded"])
    out = add_b200.full_image_eval_preprocess(spec["crop"])({'image': img_d, 'label': enc})
    assert np.array_equal(out['image'].cpu().numpy(), IO[f"{name}/image"])          # bit-identical to ToTensor+Normalize+ZeroPad2d
    assert np.array_equal(out['label'].cpu().numpy().astype(np.int64), IO[f"{name}/label"])
    # fused encode + pad in one launch
    both = add_b200.encode_segmap(ids_d, pad_to=spec["crop"])
    assert torch.equal(both, out['label'])
    # batched call == per-image calls
    out2 = add_b200.full_image_eval_preprocess(spec["crop"])({'image': torch.stack([img_d, img_d.flip(0)]), 'label': torch.stack([enc, enc.flip(0)])})
    assert torch.equal(out2['image'][0], out['image']) and torch.equal(out2['label'][0], out['label'])
    # decode: int64 (argmax output) and uint8 inputs
    for lab in (enc.long(), enc):
        assert np.array_equal(add_b200.decode_segmap(lab, 'cityscapes'), IO[f"{name}/decoded"])
    with pytest.raises(NotImplementedError):
        add_b200.decode_segmap(enc, 'pascal')


def test_full_size_eval_crop_1025x2049():
    """The authors' real evaluation shape: a 1024x2048 Cityscapes frame padded to 1025x2049 (cityscapes.py:109-119)."""
    g = torch.Generator().manual_seed(9)
    img = torch.randint(0, 256, (1024, 2048, 3), generator=g, dtype=torch.int64).to(torch.uint8)
    ids = torch.randint(0, 34, (1024, 2048), generator=g, dtype=torch.int64).to(torch.uint8)
    out = add_b200.full_image_eval_preprocess()({'image': img.to(DEV), 'label': add_b200.encode_segmap(ids.to(DEV))})
    t, m = util.orc.full_image_eval_preprocess(img.numpy(), util.orc.encode_segmap(ids.numpy()), (1025, 2049))
    assert tuple(out['image'].shape) == (3, 1025, 2049)
    assert torch.equal(out['image'].cpu(), t) and torch.equal(out['label'].cpu().long(), m)
    assert float(out['image'][:, 1024, :].abs().max()) == 0.0 and int(out['label'][1024].min()) == 255


def test_load_checkpoint_module_prefix(tmp_path):
    """eval.py:126-140: {'epoch','state_dict','optimizer','best_pred'} with DataParallel's 'module.' prefix."""
    src = util.make_net(util.NET_CASES["searched-dense-C2"])
    ckpt = {'epoch': 7, 'state_dict': {"module." + k: v for k, v in src.state_dict().items()}, 'optimizer': {}, 'best_pred': 0.5}
    path = tmp_path / "ckpt.pth.tar"
    torch.save(ckpt, path)
    dst = add_b200.build_add("searched-dense", 2, 20, seed=3)
    epoch, best = add_b200.load_checkpoint(dst, str(path), clean_module=True)
    assert (epoch, best) == (7, 0.5)
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    dst2 = add_b200.build_add("searched-dense", 2, 20, seed=4)
    add_b200.load_checkpoint(dst2, {'state_dict': src.state_dict()})          # no prefix: auto-detected
    x, _ = util.make_input(1, 33, 65)
    o1 = dst.to(DEV)(x.to(DEV)); o2 = dst2.to(DEV)(x.to(DEV))
    assert all(torch.equal(a, b) for a, b in zip(o1, o2))
    with pytest.raises(RuntimeError):
        add_b200.load_checkpoint(dst, str(tmp_path / "missing.pth"))
