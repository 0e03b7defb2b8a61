"""Shared test helpers: golden loading and weight regeneration (no oracle / reference imports here)."""
import os

import numpy as np
import torch

import recipe

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_case(name):
    case = recipe.CASES[name]
    g = np.load(os.path.join(GOLDEN, f'idisp_{name}.npz'))
    sd = recipe.make_state_dict(recipe.stack3d_shapes(case['C']), case['seed'])
    wsum = 0
    for k in sorted(sd):
        wsum = (wsum * 31 + int(recipe.checksum(sd[k])[0])) & 0x7FFFFFFFFFFF
    assert wsum == int(g['weights_crc'][0]), 'regenerated weights differ from the ones the golden was made with'
    for k in g.files:
        if k.startswith('bn/'):
            sd[k[3:]] = torch.from_numpy(g[k])
    L, R = recipe.make_features(case['B'], case['C'], case['Hf'], case['Wf'], case['seed'])
    assert int(recipe.checksum(L)[0]) == int(g['left_crc'][0]) and int(recipe.checksum(R)[0]) == int(g['right_crc'][0])
    return case, g, sd, L, R


def _crc_of(sd):
    wsum = 0
    for k in sorted(sd):
        wsum = (wsum * 31 + int(recipe.checksum(sd[k])[0])) & 0x7FFFFFFFFFFF
    return wsum


def load_psm_case(name):
    """Whole-PSMNet fixture (image crops through the real extractor): case, golden, full 514-key state_dict, crops."""
    case = recipe.PSM_CASES[name]
    g = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    shapes = dict(recipe.stack3d_shapes(32))
    shapes.update(recipe.feature2d_shapes())
    sd = recipe.make_state_dict(shapes, case['seed'])
    assert _crc_of(sd) == int(g['weights_crc'][0]), 'regenerated weights differ from the ones the golden was made with'
    for k in g.files:
        if k.startswith('bn/'):
            sd[k[3:]] = torch.from_numpy(g[k])
    L, R = recipe.make_stereo_crops(case['R'], case['size'], case['seed'])
    assert int(recipe.checksum(L)[0]) == int(g['left_crc'][0]) and int(recipe.checksum(R)[0]) == int(g['right_crc'][0])
    return case, g, sd, L, R


def load_raw_case(name):
    """Default-initialisation fixture of the 3-D stack (stackhourglass.py:90-104)."""
    case = recipe.RAW_CASES[name]
    g = np.load(os.path.join(GOLDEN, f'{name}.npz'))
    sd = recipe.make_state_dict(recipe.stack3d_shapes(case['C']), case['seed'], raw=True)
    assert _crc_of(sd) == int(g['weights_crc'][0]), 'regenerated weights differ from the ones the golden was made with'
    L, R = recipe.make_features(case['B'], case['C'], case['Hf'], case['Wf'], case['seed'])
    assert int(recipe.checksum(L)[0]) == int(g['left_crc'][0]) and int(recipe.checksum(R)[0]) == int(g['right_crc'][0])
    return case, g, sd, L, R


def make_full_psmnet(case, sd, precision='auto', device='cuda'):
    """Product PSMNet WITH its feature extractor (the drop-in module DispRCNN3D builds, disprcnn3d.py:21-33)."""
    from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet
    m = PSMNet(case['maxdisp'], case['mindisp'], precision=precision)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device).eval()


def make_psmnet(case, sd, precision='fp32', device='cuda'):
    """Product PSMNet for a feature-input config, loaded with the golden's weights."""
    import torch.nn as nn
    from disprcnn_b200.modeling.psmnet.stackhourglass import PSMNet
    m = PSMNet(case['maxdisp'], case['mindisp'], feature_channels=case['C'], precision=precision)
    m.feature_extraction = nn.Identity()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith('num_batches_tracked') for k in missing), (missing, unexpected)
    return m.to(device).eval()
