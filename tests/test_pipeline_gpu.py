"""GPU: the reference-facing predict path (segment / EnsemblePipeline / src/predict.py main) against
the oracle restatement of /root/reference/src/predict.py on the same synthetic checkpoints + frames."""
import json
import os

import cv2

import numpy as np
import pytest
import torch
from PIL import Image

from oct_segmentation_b200 import predict as pred
from oct_segmentation_b200 import prepost as P
from oct_segmentation_b200.model import OCTSegmentationModel
from oct_segmentation_b200.pipeline import EnsemblePipeline
from oracle import model_ref, synth
from oracle import prepost_ref as R

pytestmark = pytest.mark.gpu

CLASSES = ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']
SMALL = {'LM': 128, 'FC_LC': 160, 'VV': 160}


@pytest.fixture(scope='module')
def model_pairs():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    refs, ours = {}, {}
    for key, S in SMALL.items():
        r = synth.make_model(key, calib_size=S, calib_frames=2)
        cfg = dict(synth.MODEL_CONFIGS[key], input_size=S)
        o = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                 in_channels=3, classes=cfg['classes'], encoder_weights=None)
        o.load_state_dict(r.state_dict(), strict=True)
        refs[key] = (r.cuda().eval(), cfg)
        ours[key] = (o.cuda().eval(), cfg)
    return refs, ours


def test_segment_matches_reference_segment(model_pairs):
    refs, ours = model_pairs
    out_size = [250, 250]
    frames = synth.synthetic_frames(300, 3, 250)
    images = [Image.fromarray(f) for f in frames]
    want = model_ref.segment_with_models(images, [np.zeros((250, 250, 4)) for _ in images], out_size, CLASSES, refs, 'cuda')
    quantities = []
    got = pred.segment(images, [np.zeros((250, 250, 4)) for _ in images], out_size, CLASSES, models_dir='', device='cuda',
                       batch_size=2, models=ours, quantities=quantities)
    assert len(quantities) == 3
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape == (250, 250, 4) and g.dtype == np.float64
        assert set(np.unique(g)) <= {0.0, 1.0}
        for c, name in enumerate(CLASSES):
            agree = (g[:, :, c] == w[:, :, c]).mean()
            inter = ((g[:, :, c] == 1) & (w[:, :, c] == 1)).sum()
            dice = 2 * inter / max((g[:, :, c] == 1).sum() + (w[:, :, c] == 1).sum(), 1)
            print(f'frame {i} {name}: pixel agreement {agree:.4f} dice {dice:.4f}')
            # noise-like synthetic logits: a few % of pixels sit inside the bf16 error band
            assert agree > 0.93
            # area counts are bit-exact for whatever mask the GPU produced
            assert quantities[i][name]['nnz'] == int((g[:, :, c] != 0).sum())


def test_pipeline_postprocessing_is_exact_given_the_same_planes(model_pairs):
    """With the network outputs taken from the GPU, everything after them (nearest resize, routing,
    label map, counts, radial thickness) equals the CPU oracle bit for bit."""
    _, ours = model_pairs
    Ho = 333
    pipe = EnsemblePipeline(ours, CLASSES, [Ho, Ho], 'cuda:0', 2, src_hw=(250, 250), thickness=True)
    frames = synth.synthetic_frames(310, 2, 250)
    mask, label, counts, radii = pipe.run_host(frames)
    planes = {d: pipe.nets[d].out.cpu().numpy() for d in pipe.model_dirs}           # (N, C, S, S) uint8
    for n in range(2):
        model_masks = {d: planes[d][n].transpose(1, 2, 0).astype(np.float32) for d in planes}
        want = R.route_masks(model_masks, CLASSES, [Ho, Ho], model_ref.MODELS_META)
        assert np.array_equal(mask[n], want.astype(np.uint8))
        assert np.array_equal(label[n], R.label_map(want, CLASSES))
        assert np.array_equal(counts[n], [R.area_count(want[:, :, c]) for c in range(4)])
        for c in range(4):
            assert np.array_equal(radii[n, c], R.radial_radii((want[:, :, c] * 255).astype(np.uint8)))
    # input of each network == cv2 preprocessing of the frames (the pipeline writes it stem-packed: bf16, 2x2 blocks)
    for d in pipe.model_dirs:
        S = pipe.sizes[d]
        want_in = np.stack([R.preprocess_frame(f, S) for f in frames])
        x2 = pipe.nets[d].x_s2d
        assert x2.dtype == torch.bfloat16 and (x2[..., 12:] == 0).all()
        assert np.array_equal(P.unpack_s2d(x2).float().cpu().numpy(), want_in.astype(np.float32))


def test_stream_host_equals_run_host(model_pairs):
    """The pipelined host stream (double-buffered copies under compute) returns, batch by batch, exactly what
    the synchronous run_host returns -- including a ragged last batch."""
    _, ours = model_pairs
    pipe = EnsemblePipeline(ours, CLASSES, [200, 200], 'cuda:0', 2, src_hw=(160, 160), thickness=True)
    frames = synth.synthetic_frames(330, 7, 160)
    spans = [(0, 2), (2, 4), (4, 6), (6, 7)]
    want = [pipe.run_host(frames[lo:hi]) for lo, hi in spans]
    got = list(pipe.stream_host(frames[lo:hi] for lo, hi in spans))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for a, b in zip(g, w):
            assert np.array_equal(a, b)


def test_predict_main_end_to_end(tmp_path, model_pairs):
    """src/predict.py main(): models_dir with config.json + weights.ckpt, PNG inputs, PNG + JSON outputs."""
    refs, _ = model_pairs
    models_dir, data_dir, save_dir = tmp_path / 'models', tmp_path / 'in', tmp_path / 'out'
    data_dir.mkdir()
    for key, (r, cfg) in refs.items():
        (models_dir / key).mkdir(parents=True)
        synth.save_checkpoint(r.cpu(), str(models_dir / key / 'weights.ckpt'))
        json.dump(cfg, open(models_dir / key / 'config.json', 'w'))
        r.cuda()
    for i, f in enumerate(synth.synthetic_frames(320, 2, 200)):
        Image.fromarray(f).save(data_dir / f'frame_{i}.png')
    from oct_segmentation_b200 import config
    cfg = config.compose(os.path.join(os.path.dirname(os.path.dirname(__file__)), 'configs'), 'predict',
                         [f'data_dir={data_dir}', f'models_dir={models_dir}', f'save_dir={save_dir}',
                          'output_size=[256,256]', 'device=cuda', 'batch_size=2'])
    pred.main(cfg)
    for i in range(2):
        m = np.array(Image.open(save_dir / f'frame_{i}_mask.png'))
        assert m.shape == (256, 256, 3)
        colors = {tuple(c) for c in np.unique(m.reshape(-1, 3), axis=0)}
        assert colors <= {(128, 128, 128), (228, 30, 199), (123, 171, 226), (125, 227, 127), (208, 2, 27)}
        assert (save_dir / f'frame_{i}_overlay.png').exists()
    q = json.load(open(save_dir / 'quantities.json'))
    assert set(q) == {'frame_0', 'frame_1'} and set(q['frame_0']) == set(CLASSES)


def test_save_results_writes_the_reference_pngs(tmp_path):
    """predict.save_results (same arguments as src/data/utils.py:195) on the frames + float64 masks the
    reference's own save_results was run on (tests/golden/make_golden.py): both PNGs identical."""
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'overlay_ref.npz'))
    for i in range(int(d['n'])):
        classes = [str(c) for c in d[f'classes{i}']]
        pred.save_results([Image.fromarray(d[f'frame{i}'])] * 3, [d[f'mask{i}'].astype(np.float64)] * 3,
                          ['a', 'b', 'c'], classes, str(tmp_path), batch_size=2)
        for name in 'abc':
            assert np.array_equal(np.array(Image.open(tmp_path / f'{name}_overlay.png')), d[f'overlay{i}'])
            assert np.array_equal(np.array(Image.open(tmp_path / f'{name}_mask.png')), d[f'colormask{i}'])


def test_opt_in_fold_averaging(model_pairs):
    """K-way probability averaging (opt-in; a list of folds for a model_dir): the VV planes equal the float64 oracle
    average of the folds' own fp32 logits outside the rounding band, every other class is untouched, and a
    one-element list is the plain routing path."""
    _, ours = model_pairs
    vv_a, cfg = ours['VV']
    vv_b = OCTSegmentationModel(arch=cfg['architecture'], encoder_name=cfg['encoder'], model_name=cfg['model_name'],
                                in_channels=3, classes=cfg['classes'], encoder_weights=None)
    vv_b.load_state_dict(vv_a.state_dict(), strict=True)
    with torch.no_grad():                      # a second "fold": same architecture, different head
        vv_b.model.segmentation_head[0].weight.mul_(-0.7)
        vv_b.model.segmentation_head[0].bias.add_(0.3)
    vv_b.model.invalidate()
    vv_b = vv_b.cuda().eval()
    Ho = 250
    frames = synth.synthetic_frames(330, 2, 250)
    plain = EnsemblePipeline(ours, CLASSES, [Ho, Ho], 'cuda:0', 2, src_hw=(250, 250))
    want_mask = plain.run_host(frames)[0]
    one = EnsemblePipeline(dict(ours, VV=[(vv_a, cfg)]), CLASSES, [Ho, Ho], 'cuda:0', 2, src_hw=(250, 250))
    assert np.array_equal(one.run_host(frames)[0], want_mask)
    two = EnsemblePipeline(dict(ours, VV=[(vv_a, cfg), (vv_b, cfg)]), CLASSES, [Ho, Ho], 'cuda:0', 2, src_hw=(250, 250))
    mask, _, counts, _ = two.run_host(frames)
    assert np.array_equal(mask[..., :3], want_mask[..., :3])
    logits = [net.out.cpu().numpy() for net in two.fold_nets['VV']]
    assert logits[0].dtype == np.float32 and not np.array_equal(logits[0], logits[1])
    want, margin = R.fold_average_threshold(logits)
    got = two.fold_planes['VV'].cpu().numpy()
    sure = margin > 1e-6
    assert sure.mean() > 0.999 and np.array_equal(got[sure], want[sure])
    S = SMALL['VV']
    iy, ix = R.nearest_index(S, Ho), R.nearest_index(S, Ho)
    assert np.array_equal(mask[..., 3], got[:, 0][:, iy][:, :, ix])
    assert np.array_equal(counts[:, 3], (mask[..., 3] != 0).sum(axis=(1, 2)))


def test_streaming_main_equals_hold_everything_main(tmp_path, model_pairs):
    """`stream_chunk` (chunks decoded / segmented / encoded in overlap, bounded host memory) writes exactly the
    files of the reference-shaped flow: same PNG pixels, same quantities.json; ragged last chunk and last batch."""
    refs, _ = model_pairs
    models_dir, data_dir = tmp_path / 'models', tmp_path / 'in'
    data_dir.mkdir()
    for key, (r, cfg) in refs.items():
        (models_dir / key).mkdir(parents=True)
        synth.save_checkpoint(r.cpu(), str(models_dir / key / 'weights.ckpt'))
        json.dump(cfg, open(models_dir / key / 'config.json', 'w'))
        r.cuda()
    for i, f in enumerate(synth.synthetic_frames(340, 5, 200)):
        Image.fromarray(f).save(data_dir / f'frame_{i}.png')
    from oct_segmentation_b200 import config
    outs = {}
    for mode, extra in (('hold', []), ('stream', ['stream_chunk=2', 'io_workers=3'])):
        save_dir = tmp_path / mode
        cfg = config.compose(os.path.join(os.path.dirname(os.path.dirname(__file__)), 'configs'), 'predict',
                             [f'data_dir={data_dir}', f'models_dir={models_dir}', f'save_dir={save_dir}',
                              'output_size=[256,256]', 'device=cuda', 'batch_size=2'] + extra)
        pred.main(cfg)
        outs[mode] = save_dir
    assert sorted(os.listdir(outs['hold'])) == sorted(os.listdir(outs['stream']))
    assert len(os.listdir(outs['hold'])) == 12          # 5 x (mask, overlay) + quantities.json + objects.json
    for name in os.listdir(outs['hold']):
        if name.endswith('.png'):
            assert np.array_equal(np.array(Image.open(outs['hold'] / name)), np.array(Image.open(outs['stream'] / name))), name
    for name in ('quantities.json', 'objects.json'):
        assert json.load(open(outs['hold'] / name)) == json.load(open(outs['stream'] / name))
    obj = json.load(open(outs['hold'] / 'objects.json'))
    assert obj['images'] == [f'frame_{i}' for i in range(5)] and set(obj['objects']) == set(CLASSES)


def test_pipeline_contour_quantities(model_pairs):
    """contour=True: the 5th result carries the largest outer border per class; the table's contour thickness equals
    calculate_thickness_contour (analysis.py:21-57, 202-207) on the mask the GPU produced; stream_host == run_host."""
    _, ours = model_pairs
    Ho = 250
    frames = synth.synthetic_frames(350, 3, 250)
    pipe = EnsemblePipeline(ours, CLASSES, [Ho, Ho], 'cuda:0', 2, src_hw=(250, 250), thickness=True, contour=True)
    spans = [(0, 2), (2, 3)]
    want = [pipe.run_host(frames[lo:hi]) for lo, hi in spans]
    got = list(pipe.stream_host(frames[lo:hi] for lo, hi in spans))
    ratio = P.dicom_ratio(Ho)
    for g, w in zip(got, want):
        assert len(g) == len(w) == 5
        for a, b in zip(g[:4], w[:4]):
            assert np.array_equal(a, b)
        mask, _, counts, radii, contours = g
        for a, b in zip(contours, w[4]):
            assert a.shape == b.shape
        rows = P.quantities_from_counts(counts, Ho, Ho, ratio, radii, contours)
        rows_w = P.quantities_from_counts(w[2], Ho, Ho, ratio, w[3], w[4])
        assert rows == rows_w
        for n, row in enumerate(rows):
            for c, name in enumerate(CLASSES):
                ch = np.ascontiguousarray(mask[n, :, :, c])
                if row[name]['present']:
                    t = R.thickness_contour(ch)
                    assert row[name]['contour_thickness_mean'] == t['median'] / ratio
                    assert row[name]['contour_thickness_min'] == t['min'] / ratio
                else:
                    assert 'contour_thickness_mean' not in row[name]


def test_dicom_volume_analysis_matches_reference_formulas():
    """analysis.analyse_volume = get_analysis's data path (src/app/tools/analysis.py:139-213) with its
    `TODO: run inference` filled in by the GPU ensemble.  A synthetic 16-bit grayscale volume (a gap of empty slices
    in the middle) -> per-slice cv2 min-max normalise + BGR2RGB (the reference's two calls) -> segment() -> objects
    table.  Checked against the oracle's restatement of the same formulas applied to the masks the pipeline produced:
    slices, object ids, area = sqrt(nnz // ratio) with the DICOM's ratio, contour thickness,
    and the base64 PNG of every present class mask."""
    import base64
    from io import BytesIO
    from PIL import Image
    from oct_segmentation_b200 import analysis, predict as PR, synthetic
    from oracle import prepost_ref as R
    dev = torch.device('cuda')
    models = synthetic.random_models(dev, input_size=128)
    n, S = 6, 160
    vol = np.stack([synthetic.synthetic_frame(900 + i, S)[..., 0].astype(np.uint16) * 200 for i in range(n)])
    vol[2:4] = 0                                                     # empty slices: every class absent there
    out = [250, 250]
    data = analysis.analyse_volume(vol, models, output_size=out, batch_size=4)
    assert data['ratio'] == int(S * 150 // 1000) and data['images'] == [f'{i + 1:03d}' for i in range(n)]
    # the masks, recomputed through the same public call on the same normalised frames
    images = [Image.fromarray(analysis.normalise_slice(vol[i])).resize(tuple(out)) for i in range(n)]
    assert np.array_equal(np.asarray(analysis.normalise_slice(vol[0]))[..., 0],
                          cv2.normalize(vol[0], None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_8U))
    masks = [np.zeros((out[1], out[0], 4)) for _ in range(n)]
    PR.segment(images, masks, out, analysis.CLASS_NAMES, '', 'cuda', batch_size=4, models=models)
    m8 = [(m != 0).astype(np.uint8) * 255 for m in masks]
    some = False
    for c, name in enumerate(analysis.CLASS_NAMES):
        present = [R.class_present(np.ascontiguousarray(m[:, :, c])) for m in m8]
        obj = data['objects'][name]
        assert obj['slice'] == [i for i, p in enumerate(present) if p]
        assert obj['object_id'] == R.object_ids(present)
        want = [R.frame_quantities(m, data['ratio'])[name] for m, p in zip(m8, present) if p]
        assert obj['area'] == [w['area'] for w in want]
        assert obj['thickness_mean'] == [w['thickness_mean'] for w in want]
        assert obj['thickness_min'] == [w['thickness_min'] for w in want]
        assert obj['img_name'] == [f'{i + 1:03d}' for i, p in enumerate(present) if p]
        for b64, idx in zip(obj['masks'], obj['slice']):
            assert np.array_equal(np.asarray(Image.open(BytesIO(base64.b64decode(b64)))), m8[idx][:, :, c])
        some |= len(obj['slice']) > 0
    assert some, 'no class present in any slice: the test volume is degenerate'
