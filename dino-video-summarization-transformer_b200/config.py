"""Config resolution and the decode front half of the scoring driver, so that `dino_similarity(args, ...)` can be called
with exactly the reference's arguments.

  * `load_config(args)`     - utils/parser.py:65-90 for the fields this path reads: defaults (utils/defaults.py), then the
                              YAML file `args.cfg_file`, then the `args.opts` KEY VALUE list.  fvcore / yacs are not
                              needed; unknown keys are kept (the reference sets ad-hoc keys at run time too).
  * `DecodedVideoDataset`   - datasets_custom/dino_loss_loader.py:10-92 up to (not including) view construction: reads
                              `<DATA.PATH_TO_DATA_DIR>/<mode>.csv`, decodes each video with torchvision.io.read_video and
                              yields (decoded uint8 frames [n,H,W,3], path).  Sub-sampling, normalisation, the centre crop
                              and the local / global windows happen on the device (dino_similarity.preprocess_frames,
                              dvst_forward_windows) - the reference's [2F,3,G,224,224] padded views are never built.
"""
import ast
import os


class Node:
    """Attribute bag with dict-style access (what the path needs of fvcore's CfgNode)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return self.__dict__[k]

    def __contains__(self, k):
        return k in self.__dict__

    def __repr__(self):
        return "Node(%s)" % ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items())


def default_config():
    """The defaults of utils/defaults.py for the keys the scoring path reads (line numbers of the reference file)."""
    return Node(
        TEST=Node(NUM_ENSEMBLE_VIEWS=10),                                              # :96
        MODEL=Node(NUM_CLASSES=400),                                                   # :212
        TIMESFORMER=Node(ATTENTION_TYPE='divided_space_time', PRETRAINED_MODEL=''),    # :258-259
        DATA=Node(PATH_TO_DATA_DIR="", PATH_LABEL_SEPARATOR=" ", PATH_PREFIX="", NUM_FRAMES=8, SAMPLING_RATE=8,
                  MEAN=[0.45, 0.45, 0.45], STD=[0.225, 0.225, 0.225], TRAIN_CROP_SIZE=224),   # :285-315
    )


def _merge_dict(node, d):
    for k, v in d.items():
        if isinstance(v, dict):
            sub = getattr(node, k, None)
            if not isinstance(sub, Node):
                sub = Node()
                setattr(node, k, sub)
            _merge_dict(sub, v)
        else:
            setattr(node, k, v)


def _literal(v):
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def load_config(args):
    """utils/parser.py:65-90.  `args` needs `cfg_file` (or None) and `opts` (flat KEY VALUE list or None); an object
    that already looks like a resolved config (has DATA and TIMESFORMER) is returned unchanged."""
    if hasattr(args, "DATA") and hasattr(args, "TIMESFORMER") and not getattr(args, "cfg_file", None) \
            and not getattr(args, "opts", None):
        return args
    cfg = default_config()
    cfg_file = getattr(args, "cfg_file", None)
    if cfg_file is not None:
        import yaml
        with open(cfg_file, "r") as f:
            _merge_dict(cfg, yaml.safe_load(f) or {})
    opts = getattr(args, "opts", None)
    if opts:
        if len(opts) % 2:
            raise ValueError("opts must be a flat KEY VALUE list")
        for key, val in zip(opts[0::2], opts[1::2]):
            node = cfg
            parts = key.split(".")
            for p in parts[:-1]:
                if not isinstance(getattr(node, p, None), Node):
                    setattr(node, p, Node())
                node = getattr(node, p)
            setattr(node, parts[-1], _literal(val))
    if hasattr(args, "num_shards") and hasattr(args, "shard_id"):
        cfg.NUM_SHARDS, cfg.SHARD_ID = args.num_shards, args.shard_id
    if hasattr(args, "rng_seed"):
        cfg.RNG_SEED = args.rng_seed
    if hasattr(args, "output_dir"):
        cfg.OUTPUT_DIR = args.output_dir
    return cfg


class DecodedVideoDataset:
    """The video list and decode step of DinoLossLoader (datasets_custom/dino_loss_loader.py:40-82)."""

    def __init__(self, cfg, mode="test"):
        path_to_file = os.path.join(cfg.DATA.PATH_TO_DATA_DIR, "{}.csv".format(mode))
        if not os.path.exists(path_to_file):
            raise FileNotFoundError("{} dir not found".format(path_to_file))
        self._path_to_videos = []
        self._labels = []
        with open(path_to_file, "r") as f:
            for path_label in f.read().splitlines():
                parts = path_label.split(cfg.DATA.PATH_LABEL_SEPARATOR)
                if len(parts) != 2:
                    raise ValueError(f"bad line in {path_to_file}: {path_label!r}")
                for _ in range(cfg.TEST.NUM_ENSEMBLE_VIEWS):
                    self._path_to_videos.append(os.path.join(cfg.DATA.PATH_PREFIX, parts[0]))
                    self._labels.append(int(parts[1]))
        if not self._path_to_videos:
            raise ValueError("Failed to load data split {} from {}".format(mode, path_to_file))
        print("Constructing dataloader (size: {}) from {}".format(len(self._path_to_videos), path_to_file))

    def __len__(self):
        return len(self._path_to_videos)

    def __getitem__(self, index):
        import torchvision.io as io
        read_video = getattr(io, "read_video", None)
        if read_video is None:
            raise RuntimeError("this torchvision has no io.read_video (the decoder the reference uses, "
                               "dino_loss_loader.py:81): pass `dataset=` an iterable of (decoded uint8 frames, path)")
        video, _, _ = read_video(self._path_to_videos[index], pts_unit='sec')
        return video, self._path_to_videos[index]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]
