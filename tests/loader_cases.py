"""The synthetic WeChat-shaped frame and vocabulary used to pin rank_b200.loader to the reference datasets."""
import os

import numpy as np
import pandas as pd

DENSE = ["videoplayseconds", "u_read_comment_7d_sum", "u_like_7d_sum", "u_click_avatar_7d_sum",
         "u_forward_7d_sum", "u_comment_7d_sum", "u_follow_7d_sum", "u_favorite_7d_sum",
         "i_read_comment_7d_sum", "i_like_7d_sum", "i_click_avatar_7d_sum", "i_forward_7d_sum",
         "i_comment_7d_sum", "i_follow_7d_sum", "i_favorite_7d_sum", "c_user_author_read_comment_7d_sum"]
VOCAB_FILE = {"userid": "userid.txt", "feedid": "feedid.txt", "device": "device.txt", "authorid": "authorid.txt",
              "bgm_song_id": "bgm_song_id.txt", "bgm_singer_id": "bgm_singer_id.txt", "manual_tag_list": "manual_tag_id.txt"}
N = 57


def write_vocab(path):
    os.makedirs(path, exist_ok=True)
    lines = {"userid": [f"u{i}" for i in range(30)], "feedid": [f"f{i}" for i in range(40)] + ["f3"],   # duplicated line
             "device": ["1", "2"], "authorid": [f"a{i}" for i in range(12)],
             "bgm_song_id": [f"s{i}" for i in range(9)], "manual_tag_list": ["t1", "t2", "t3"]}    # no bgm_singer_id file
    for col, ls in lines.items():
        with open(os.path.join(path, VOCAB_FILE[col]), "w") as f:
            f.write("".join(l + "\n" for l in ls))
    return path


def make_frame(history="string", seed=7):
    """history: 'string' (comma separated, as the ETL writes it) or 'list' (python lists / arrays after parquet)."""
    rng = np.random.default_rng(seed)
    pick = lambda prefix, hi, n=N: [f"{prefix}{rng.integers(0, hi)}" for _ in range(n)]
    df = pd.DataFrame({
        "userid": pick("u", 36),                       # some ids are outside the vocabulary
        "feedid": pick("f", 45),
        "device": rng.integers(1, 3, N),               # integer column: never equals a vocabulary line
        "authorid": pick("a", 12),
        "bgm_song_id": [None if i % 5 == 0 else f"s{rng.integers(0, 9)}" for i in range(N)],   # missing values
        "bgm_singer_id": pick("g", 5),                 # vocabulary file absent
        "manual_tag_list": pick("t", 5),
        "read_comment": (rng.random(N) < 0.3).astype(np.int64),
    })
    for i, c in enumerate(DENSE[:-1]):                 # the last dense column is absent from the frame
        df[c] = rng.poisson(3.0, N).astype(np.float64) if i % 2 else rng.random(N).astype(np.float32)
    lens = rng.integers(0, 9, N)
    hist = [[f"f{rng.integers(0, 45)}" for _ in range(l)] for l in lens]
    df["his_read_comment_7d_seq"] = [",".join(h) for h in hist] if history == "string" else hist
    return df
