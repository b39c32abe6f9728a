"""Vocabulary handling shared by the models.

Every reference model sizes its tables as `len(vocab file) + 1` (row 0 = unknown), re-reading
the text files in its constructor (e.g. DeepFM/deepfm.py:79-86, DCN/dcn.py:118-126).  The files
themselves are data, not code; on a box without the reference checkout `write_vocab_dir`
produces files with the same line counts so the constructors size identical tables.
"""
from __future__ import annotations

import os

# column -> vocabulary file, and the line counts of the WeChat-challenge vocabularies
# (dataset/wechat_algo_data1/vocabulary/*.txt in the reference, `wc -l`).
VOCAB_FILE = {
    "userid": "userid.txt",
    "feedid": "feedid.txt",
    "device": "device.txt",
    "authorid": "authorid.txt",
    "bgm_song_id": "bgm_song_id.txt",
    "bgm_singer_id": "bgm_singer_id.txt",
    "manual_tag_list": "manual_tag_id.txt",
}
WECHAT_VOCAB_LINES = {
    "userid": 19626,
    "feedid": 106444,
    "device": 2,
    "authorid": 18789,
    "bgm_song_id": 25159,
    "bgm_singer_id": 17500,
    "manual_tag_list": 350,
}


def count_vocab_lines(vocab_dir: str, filename: str) -> int:
    """Lines of one vocabulary file; a missing file is an empty vocabulary (reference
    `_load_vocabulary` returns [] then, e.g. DeepFM/deepfm.py:114-119)."""
    path = os.path.join(vocab_dir, filename)
    if not os.path.exists(path):
        return 0
    with open(path, "r") as f:
        return sum(1 for _ in f)


def table_heights(vocab_dir: str, columns) -> dict:
    """{column: vocabulary size + 1} in the order given."""
    return {c: count_vocab_lines(vocab_dir, VOCAB_FILE[c]) + 1 for c in columns}


def write_vocab_dir(path: str, lines: dict | None = None) -> str:
    """Create vocabulary files with the given line counts (default: the WeChat sizes)."""
    lines = dict(WECHAT_VOCAB_LINES if lines is None else lines)
    os.makedirs(path, exist_ok=True)
    for col, n in lines.items():
        with open(os.path.join(path, VOCAB_FILE[col]), "w") as f:
            f.write("".join(f"{col}_{i}\n" for i in range(1, n + 1)))
    return path
