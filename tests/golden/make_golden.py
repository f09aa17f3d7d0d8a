"""Generates tests/golden/* from the reference's own fixtures (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

  masks_app_demo.npz   : 12 of the 186 4-channel {0,255} masks in data/app/demo/mask, bit-packed,
                         + per-class non-zero counts of ALL 186 masks
  colorize_pairs.npz   : data/visualization mask -> mask_color pairs (pins the priority merge rule,
                         src/data/convert_int_to_cv.py:96-108 == src/data/utils.py:231-233)
  demo_frame_small.npz : one real OCT frame (data/demo/input/001_1_007.png) downsampled to 250x250
                         by plain decimation, used as a realistic pre-processing input
  overlay_ref.npz      : <name>_overlay.png / <name>_mask.png written BY THE REFERENCE'S OWN save_results
                         (src/data/utils.py:195-235) for small frames + masks (shapes touching every border)
"""
import ast
import glob
import math
import os

import numpy as np
from PIL import Image

REF = '/root/reference'
OUT = os.path.dirname(os.path.abspath(__file__))


def reference_functions(path, names, extra=None):
    """Exec selected top-level functions of a reference source file (its module-level imports of
    gradio / pydicom / hydra ... are not installable here) with the real cv2 / numpy / math."""
    import cv2
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'cv2': cv2, 'np': np, 'math': math, 'Dict': dict, 'Any': object, 'List': list, 'Tuple': tuple,
          'Union': object, 'Image': Image}
    ns.update(extra or {})
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, 'exec'), ns)
    return [ns[n] for n in names]


def main():
    paths = sorted(glob.glob(f'{REF}/data/app/demo/mask/*.tiff'))
    counts = np.zeros((len(paths), 4), np.int64)
    shapes = []
    keep_idx = list(range(0, len(paths), 16))[:12]
    packed, names = [], []
    for i, p in enumerate(paths):
        m = np.array(Image.open(p))
        assert m.ndim == 3 and m.shape[2] == 4 and set(np.unique(m)) <= {0, 255}
        counts[i] = [(m[:, :, c] != 0).sum() for c in range(4)]
        shapes.append(m.shape[:2])
        if i in keep_idx:
            packed.append(np.packbits(m != 0, axis=None))
            names.append(os.path.basename(p))
    np.savez_compressed(f'{OUT}/masks_app_demo.npz', counts=counts, names=np.array(names), keep_idx=np.array(keep_idx),
                        packed=np.stack(packed), shape=np.array(shapes[0]),
                        all_names=np.array([os.path.basename(p) for p in paths]))

    masks, colors, nm = [], [], []
    for p in sorted(glob.glob(f'{REF}/data/visualization/mask/*.tiff')):
        stem = os.path.basename(p).split('.')[0]
        cp = f'{REF}/data/visualization/mask_color/{stem}.tiff'
        if not os.path.exists(cp):
            continue
        m = np.array(Image.open(p))
        c = np.array(Image.open(cp).convert('RGB'))
        masks.append(np.packbits(m != 0, axis=None))
        colors.append(c)
        nm.append(stem)
    np.savez_compressed(f'{OUT}/colorize_pairs.npz', packed=np.stack(masks), colors=np.stack(colors), names=np.array(nm),
                        shape=np.array(colors[0].shape[:2]))

    img = np.array(Image.open(f'{REF}/data/demo/input/001_1_007.png').convert('RGB'))
    small = img[::3, ::3].copy()
    (preprocessing_img,) = reference_functions(f'{REF}/src/data/utils.py', ['preprocessing_img'])
    np.savez_compressed(f'{OUT}/demo_frame_small.npz', rgb=small,
                        pre96=preprocessing_img(Image.fromarray(small), 96),
                        pre160=preprocessing_img(Image.fromarray(small), 160),
                        pre125=preprocessing_img(Image.fromarray(small), 125))

    # quantities computed BY THE REFERENCE'S OWN FUNCTIONS (src/app/tools/analysis.py) on the kept masks
    contour_fn, radial_fn = reference_functions(f'{REF}/src/app/tools/analysis.py',
                                                ['calculate_thickness_contour', 'calculate_object_thickness'])
    ratio = int(750 * 150 // 1000)                       # analysis.py:155 for the 750-pixel demo frames
    q = np.zeros((len(keep_idx), 4, 8))                  # present, nnz, area, c_median, c_min, r_median, r_min, r_max
    radii_sets = []
    for k, i in enumerate(keep_idx):
        m = np.array(Image.open(paths[i]))
        for c in range(4):
            ch = np.ascontiguousarray(m[:, :, c])
            present = np.unique(ch).shape[0] == 2        # analysis.py:189
            nnz = len(np.nonzero(ch)[0])
            area = pow(nnz // ratio, 0.5)                # analysis.py:199-200
            ct = contour_fn(ch)
            rt = radial_fn(ch)
            q[k, c] = [present, nnz, area, ct['median'], ct['min'], rt['median'], rt['min'], rt['max']]
            rr = np.full(360, -1, np.int32)
            rr[:len(rt['all_measurements'])] = rt['all_measurements']      # hits only, in angle order
            radii_sets.append(rr)
    np.savez_compressed(f'{OUT}/quantities_ref.npz', q=q, ratio=ratio, keep_idx=np.array(keep_idx),
                        radii_hits=np.stack(radii_sets).reshape(len(keep_idx), 4, 360))
    # overlay cosmetics: run the reference's save_results itself
    import tempfile
    (union_fn,) = reference_functions(f'{REF}/src/models/smp/utils.py', ['get_img_mask_union_pil'])
    consts = {}
    tree = ast.parse(open(f'{REF}/src/data/utils.py').read())
    for node in tree.body:
        if isinstance(node, ast.Assign) and any(getattr(t, 'id', '') in ('CLASS_MAP', 'CLASS_COLORS_RGB', 'CLASS_IDS')
                                                for t in node.targets):
            exec(compile(ast.Module([node], []), 'utils.py', 'exec'), consts)
    consts.update({'get_img_mask_union_pil': union_fn, 'tqdm': lambda it, **k: it})
    (save_results,) = reference_functions(f'{REF}/src/data/utils.py', ['save_results'], extra=consts)
    rng = np.random.default_rng(7)
    cases = []
    for (H, W), classes in (((96, 128), ['Lumen', 'Fibrous cap', 'Lipid core', 'Vasa vasorum']),
                            ((50, 37), ['Vasa vasorum', 'Lumen']), ((64, 64), ['Lipid core', 'Fibrous cap', 'Lumen'])):
        frame = small[:H, :W].copy() if small.shape[0] >= H and small.shape[1] >= W else rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        yy, xx = np.mgrid[:H, :W]
        mask = np.zeros((H, W, 4))
        mask[:, :, 0] = (yy - H / 2) ** 2 + (xx - W / 2) ** 2 < (H / 3) ** 2
        mask[:, :, 1] = (yy - H / 3) ** 2 / 50 + (xx - 2 * W / 3) ** 2 / 300 < 1
        mask[:, :, 2] = rng.random((H, W)) > 0.9
        mask[:, :, 3] = (abs(yy - H * 0.7) < 3) & (xx > 5)
        mask[0:4, 0:9, 0] = 1
        mask[H - 3:, W - 6:, 1] = 1
        mask[:, 0, 2] = 1
        mask[0, :, 3] = 1
        d = tempfile.mkdtemp()
        save_results([Image.fromarray(frame)], [mask.copy()], ['x'], classes, d)
        cases.append((frame, mask.astype(np.uint8), classes, np.array(Image.open(f'{d}/x_overlay.png')),
                      np.array(Image.open(f'{d}/x_mask.png'))))
    np.savez_compressed(f'{OUT}/overlay_ref.npz', n=len(cases),
                        **{f'frame{i}': c[0] for i, c in enumerate(cases)}, **{f'mask{i}': c[1] for i, c in enumerate(cases)},
                        **{f'classes{i}': np.array(c[2]) for i, c in enumerate(cases)},
                        **{f'overlay{i}': c[3] for i, c in enumerate(cases)}, **{f'colormask{i}': c[4] for i, c in enumerate(cases)})
    print('wrote', os.listdir(OUT))


if __name__ == '__main__':
    main()
