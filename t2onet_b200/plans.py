"""The planner's on-disk records and their reader: what preprocess/gen_greedy_seqs_FiveK.py:66-83 writes per image pair
(`<save_dir>/<phase><i>/<i:05d>.json` = {'request', 'init distance', 'operation sequence'} plus input.jpg / target.jpg /
edit<k>.jpg) and what datasets/FiveKdataset.py:86-120 (`FiveKAct.get_act`) reads back for training: the top sequence,
truncated where the distance stops improving by more than 1 % of the initial distance (`analyze_traj`, :54-64), as
operator-vocabulary ids and normalised parameter rows.

`replay` is the lossless alternative to the edit<k>.jpg files: the intermediate images of a stored sequence are
re-executed from the input on the GPU (one per-step launch each) instead of being read back JPEG-quantised."""
import json
import os

import numpy as np
import torch

ACTIONS = ['brightness', 'contrast', 'saturation', 'color', 'inpaint', 'tone', 'sharpness', 'white']
ACT2PN = {'brightness': 1, 'contrast': 1, 'saturation': 1, 'color': 24, 'inpaint': 0, 'tone': 8, 'sharpness': 1, 'white': 0}
OP_VOCAB_OFFSET = 3          # operator-vocabulary id = Executor index + 3 (<NONE> 0, <START> 1, <END> 2; models/actor.py:165)


def tensor2img(tensor):
    """utils/visual_utils.py:50-58: (1,3,H,W) float RGB in [0,1] -> (H,W,3) uint8 BGR (truncating cast, as astype does)."""
    out = tensor.squeeze(0).permute(1, 2, 0) * 255
    return out.cpu().numpy().astype(np.uint8)[:, :, ::-1]


def img2tensor(img):
    """utils/visual_utils.py:61-70"""
    return (torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1].transpose(2, 0, 1))) / 255).unsqueeze(0)


def analyze_traj(seq):
    """datasets/FiveKdataset.py:54-64: number of leading steps that each improve the distance by > 1 % of the initial
    distance (at least 1)."""
    seq = np.array(seq)
    diffs = seq[:-1] - seq[1:]
    over_shot = diffs / seq[0]
    stop = np.where((over_shot > 0.01) == False)[0]          # noqa: E712  (the reference's expression)
    trunc_len = int(stop[0]) if len(stop) else len(over_shot)
    return trunc_len if trunc_len != 0 else 1


def plan_record(request, init_dist, act_seqs):
    """The JSON body of preprocess/gen_greedy_seqs_FiveK.py:74."""
    return {'request': request, 'init distance': float(init_dist),
            'operation sequence': [[[a[0], [float(v) for v in a[1]], float(a[2])] for a in seq] for seq in act_seqs]}


def write_plan(save_dir, phase, i, request, input_img, target_img, act_seqs, img_seqs, init_dist, write_images=True):
    """preprocess/gen_greedy_seqs_FiveK.py:66-83 for item i; returns the item directory."""
    item_dir = os.path.join(save_dir, '{}{}'.format(phase, i))
    os.makedirs(item_dir, exist_ok=True)
    with open(os.path.join(item_dir, '{:05d}.json'.format(i)), 'w') as f:
        json.dump(plan_record(request, init_dist, act_seqs), f)
    if write_images:
        import cv2
        cv2.imwrite(os.path.join(item_dir, 'input.jpg'), tensor2img(input_img))
        cv2.imwrite(os.path.join(item_dir, 'target.jpg'), tensor2img(target_img))
        if len(img_seqs) > 0:
            for idx, img in enumerate(img_seqs[0]):
                cv2.imwrite(os.path.join(item_dir, 'edit{}.jpg'.format(idx)), tensor2img(img))
    return item_dir


def encode_plan(record, op_max_len=5):
    """FiveKAct.get_act without the image loading (datasets/FiveKdataset.py:86-113):
    -> (op_seq (op_max_len + 2,) int, params (op_max_len, 24) float32, trunc_len, the truncated top sequence)."""
    init_dist = record['init distance']
    seq = record['operation sequence'][0]                   # the top sequence
    seq_dist = [v[2] for v in seq]
    seq_dist.insert(0, init_dist)
    trunc_len = min(analyze_traj(seq_dist), op_max_len)
    seq = seq[:trunc_len]
    params = np.zeros((op_max_len, 24), dtype=np.float32)
    op_seq = np.zeros(op_max_len + 2, dtype=int)
    i = -1
    for i, act in enumerate(seq):
        op_seq[i + 1] = ACTIONS.index(act[0]) + OP_VOCAB_OFFSET
        param_num = ACT2PN[act[0]]
        if act[0] == 'color' or act[0] == 'tone':
            max_abs = np.abs(np.array(act[1])).max()
            params[i, :param_num] = np.array(act[1]) / max_abs
        elif np.abs(act[1][0]) > 5:                         # a runaway scalar fit: predict 0
            params[i, :param_num] = np.array([0])
        else:
            params[i, :param_num] = np.array(act[1])
    op_seq[0] = 1                                           # <START>
    op_seq[i + 2] = 2                                       # <END>
    return op_seq, params, trunc_len, seq


def read_plan(act_dir, phase, item, op_max_len=5):
    """Load and encode item's record; see encode_plan."""
    item_dir = os.path.join(act_dir, '{}{}'.format(phase, item))
    with open(os.path.join(item_dir, '{:05d}.json'.format(item)), 'r') as f:
        record = json.load(f)
    return encode_plan(record, op_max_len)


def load_train_img(img_path, img_size):
    """utils/visual_utils.py:6-14"""
    import cv2
    img = cv2.resize(cv2.imread(img_path), (img_size, img_size))
    return torch.from_numpy(np.ascontiguousarray(img[:, :, ::-1].astype(np.float32).transpose(2, 0, 1))) / 255


def replay(I_0, seq, executor):
    """Intermediate images of a stored sequence, re-executed on the GPU: [op_1(I_0), op_2(op_1(I_0)), ...] with the
    stored (un-normalised) parameters -- bit-identical to the planner's own I_out for sequences it returned, where
    FiveKAct reads JPEG-quantised edit<k>.jpg files."""
    img, outs = I_0, []
    for act in seq:
        param = torch.tensor([list(act[1])], dtype=torch.float32, device=I_0.device)
        img, _ = executor.execute(img, ACTIONS.index(act[0]), None, specified_param=param)
        outs.append(img)
    return outs
