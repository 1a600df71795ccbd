"""Run directory and scalar log in the reference's format (mirror of util/log.py:11-70 and the add_scalar calls of
rl/algos/ppo.py:486-499), so that the reference's tooling — TensorBoard on the run directory, `parse_previous`
(util/log.py:74-91) and apex.py eval's `pickle.load(experiment.pkl)` (apex.py:257-280) — reads runs made here.

create_logger(args): same directory rule (logdir/env_name/run_name, or logdir/env_name/<md5 of the sorted hyper-parameters,
6 hex>-seed<seed>), same experiment.info (one `key: value` line per sorted hyper-parameter, seed / logdir / run_name
removed) and experiment.pkl (the pickled args namespace).  The returned object has `.dir` and `add_scalar(tag, value, step)`
like the torch.utils.tensorboard.SummaryWriter the reference creates; TensorBoard itself is not a dependency: ScalarWriter
emits the TFRecord / Event wire format directly (length, masked CRC-32C, payload, masked CRC-32C; Event protobuf with
wall_time, step and one Summary.Value{tag, simple_value}).
"""
import hashlib
import os
import pickle
import socket
import struct
import time
from collections import OrderedDict

PPO_SCALARS = ("Test/Return", "Train/Return", "Train/Mean Eplen", "Train/Mean KL Div", "Train/Mean Entropy", "Misc/Critic Loss",
               "Misc/Actor Loss", "Misc/Mirror Loss", "Misc/Timesteps", "Misc/Sample Times", "Misc/Optimize Times",
               "Misc/Evaluation Times", "Misc/Termination Threshold")  # rl/algos/ppo.py:486-499

_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def _event(wall_time, step=None, file_version=None, tag=None, value=None):
    ev = _field(1, 1, struct.pack("<d", wall_time))
    if step is not None:
        ev += _field(2, 0, _varint(step & 0xFFFFFFFFFFFFFFFF))
    if file_version is not None:
        fv = file_version.encode()
        ev += _field(3, 2, _varint(len(fv)) + fv)
    if tag is not None:
        t = tag.encode()
        val = _field(1, 2, _varint(len(t)) + t) + _field(2, 5, struct.pack("<f", float(value)))  # Summary.Value{tag, simple_value}
        summ = _field(1, 2, _varint(len(val)) + val)                                               # Summary{value}
        ev += _field(5, 2, _varint(len(summ)) + summ)                                              # Event.summary
    return ev


def _record(data):
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", _masked(head)) + data + struct.pack("<I", _masked(data))


class ScalarWriter:
    """add_scalar / flush / close of a SummaryWriter, writing events.out.tfevents.<time>.<host> under `dir`."""

    def __init__(self, log_dir):
        self.dir = log_dir
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self.f = open(self.path, "wb")
        self.f.write(_record(_event(time.time(), file_version="brain.Event:2")))
        self.f.flush()

    def add_scalar(self, tag, scalar_value, global_step=None, walltime=None):
        self.f.write(_record(_event(time.time() if walltime is None else walltime, step=0 if global_step is None else int(global_step),
                                    tag=tag, value=scalar_value)))

    def flush(self):
        self.f.flush()

    def close(self):
        self.f.close()


def read_scalars(path):
    """Parse an event file back into [(step, tag, value)], checking every CRC (tests; also reads files TensorBoard wrote as long
    as they only hold simple_value scalars)."""
    out, data = [], open(path, "rb").read()
    pos = 0

    def varint(buf, i):
        n = s = 0
        while True:
            b = buf[i]
            i += 1
            n |= (b & 0x7F) << s
            s += 7
            if not b & 0x80:
                return n, i

    def fields(buf):
        i = 0
        while i < len(buf):
            key, i = varint(buf, i)
            num, wire = key >> 3, key & 7
            if wire == 0:
                v, i = varint(buf, i)
            elif wire == 1:
                v, i = buf[i:i + 8], i + 8
            elif wire == 5:
                v, i = buf[i:i + 4], i + 4
            elif wire == 2:
                n, i = varint(buf, i)
                v, i = buf[i:i + n], i + n
            else:
                raise ValueError("unsupported wire type")
            yield num, wire, v
    while pos < len(data):
        head = data[pos:pos + 8]
        (n,) = struct.unpack("<Q", head)
        assert struct.unpack("<I", data[pos + 8:pos + 12])[0] == _masked(head), "length CRC"
        body = data[pos + 12:pos + 12 + n]
        assert struct.unpack("<I", data[pos + 12 + n:pos + 16 + n])[0] == _masked(body), "data CRC"
        pos += 16 + n
        step, summ = 0, None
        for num, wire, v in fields(body):
            if num == 2:
                step = v
            elif num == 5:
                summ = v
        if summ is not None:
            for num, wire, v in fields(summ):
                tag = val = None
                for n2, w2, v2 in fields(v):
                    if n2 == 1:
                        tag = v2.decode()
                    elif n2 == 2 and w2 == 5:
                        (val,) = struct.unpack("<f", v2)
                out.append((step, tag, val))
    return out


def run_directory(args):
    """The directory rule of util/log.py:22-50 (without the `previous` continuation branch, which is CLI logic)."""
    arg_dict = OrderedDict(sorted(dict(vars(args)).items(), key=lambda t: t[0]))
    for key in ("seed", "logdir", "env_name"):
        assert key in arg_dict, f"You must provide a '{key}' key in your command line arguments"
    run_name = arg_dict.pop("run_name", None)
    seed, logdir, env_name = str(arg_dict.pop("seed")), str(arg_dict.pop("logdir")), str(arg_dict["env_name"])
    if run_name is not None:
        return os.path.join(logdir, env_name, run_name), arg_dict
    arg_hash = hashlib.md5(str(arg_dict).encode("ascii")).hexdigest()[0:6] + "-seed" + seed
    return os.path.join(logdir, env_name, arg_hash), arg_dict


def create_logger(args):
    output_dir, arg_dict = run_directory(args)
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "experiment.pkl"), "wb") as f:
        pickle.dump(args, f)
    with open(os.path.join(output_dir, "experiment.info"), "w") as f:
        for key, val in arg_dict.items():
            f.write("%s: %s\n" % (key, val))
    return ScalarWriter(output_dir)


def log_ppo_iteration(logger, itr, test_return, train_return, mean_eplen, kl, entropy, critic_loss, actor_loss, mirror_loss, timesteps,
                      sample_time, optimize_time, eval_time, term_thresh=0.0):
    """The thirteen scalars of one PPO iteration under the reference's tags (rl/algos/ppo.py:486-499)."""
    vals = (test_return, train_return, mean_eplen, kl, entropy, critic_loss, actor_loss, mirror_loss, timesteps, sample_time,
            optimize_time, eval_time, term_thresh)
    for tag, v in zip(PPO_SCALARS, vals):
        logger.add_scalar(tag, v, itr)
    logger.flush()
