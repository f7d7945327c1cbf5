"""Build tests/golden/movielens_matrices.npz from the reference's movielens data files (container only).

BASELINE config C3 (examples/movielens_completion.py:20-86): users x movies ratings (mostly unknown), movies x
genres, movies x actors.  The example picks movies / actors through set iteration order and hides ratings with an
unseeded RNG; here the selection is deterministic (sorted ids, first 1000 movies / actors, RandomState(0) for the
10 % of known ratings that are hidden) so CPU oracle and GPU engine see identical inputs.  Only derived index /
value lists are stored."""
import csv
import gzip
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = "/root/reference/skfusion/datasets/data/movielens"


def main():
    ratings = {}
    with gzip.open(os.path.join(DATA, "ratings.csv.gz"), "rt", encoding="utf-8") as f:
        f.readline()
        for line in f:
            u, m, r = line.strip().split(",")[:3]
            ratings.setdefault(int(u), {})[int(m)] = float(r)
    genres, actors = {}, {}
    with gzip.open(os.path.join(DATA, "movies.csv.gz"), "rt", encoding="utf-8") as f:
        f.readline()
        for row in csv.reader(f):
            genres[int(row[0])] = row[2].split("|")
    with gzip.open(os.path.join(DATA, "actors.csv.gz"), "rt", encoding="utf-8") as f:
        f.readline()
        for row in csv.reader(f):
            actors[int(row[0])] = row[2].split("|")
    movies = sorted(set(m for val in ratings.values() for m in val))[:1000]
    movie2id = {m: i for i, m in enumerate(movies)}
    user2id = {u: i for i, u in enumerate(sorted(ratings))}
    genre2id = {g: i for i, g in enumerate(sorted(set(g for v in genres.values() for g in v)))}
    actor_list = sorted(set(a for m, v in actors.items() if m in movie2id for a in v))[:1000]
    actor2id = {a: i for i, a in enumerate(actor_list)}
    r_u, r_m, r_v = [], [], []
    for u, val in ratings.items():
        for m, r in val.items():
            if m in movie2id:
                r_u.append(user2id[u]); r_m.append(movie2id[m]); r_v.append(r)
    g_m, g_g = [], []
    for m, gs in genres.items():
        if m in movie2id:
            for g in gs:
                g_m.append(movie2id[m]); g_g.append(genre2id[g])
    a_m, a_a = [], []
    for m, as_ in actors.items():
        if m in movie2id:
            for a in as_:
                if a in actor2id:
                    a_m.append(movie2id[m]); a_a.append(actor2id[a])
    out = os.path.join(HERE, "movielens_matrices.npz")
    np.savez_compressed(out, shape=np.array([len(user2id), len(movie2id), len(genre2id), len(actor2id)]),
                        r_u=np.array(r_u, np.int32), r_m=np.array(r_m, np.int32), r_v=np.array(r_v, np.float32),
                        g_m=np.array(g_m, np.int32), g_g=np.array(g_g, np.int32),
                        a_m=np.array(a_m, np.int32), a_a=np.array(a_a, np.int32))
    print("wrote %s (%.1f KB): users=%d movies=%d genres=%d actors=%d ratings=%d" % (
        out, os.path.getsize(out) / 1024., len(user2id), len(movie2id), len(genre2id), len(actor2id), len(r_v)))


if __name__ == "__main__":
    main()
