"""Golden vectors for the metadata one-hot + StandardScaler path, produced with the objects the reference itself uses
(models/skinLesionDatasets.py:133-176): OneHotEncoder(sparse_output=False, handle_unknown='ignore') on the categorical
columns cast to str, StandardScaler on ['age', 'diameter_1', 'diameter_2'] with NaN -> -1, np.hstack((categorical,
numerical)), rounded to fp32 like the dataset's torch.tensor(metadata, dtype=torch.float32) (:56).  The frame is a seeded
synthetic PAD-UFES-20-shaped table (18 categorical + 3 numerical columns -> 85 features).  Run: python tests/golden/make_golden_metadata.py"""
import os

import numpy as np
import pandas as pd
from sklearn.preprocessing import OneHotEncoder, StandardScaler

HERE = os.path.dirname(os.path.abspath(__file__))
TRI = ["True", "False", "EMPTY"]
COLUMNS = {   # PAD-UFES-20 metadata columns after dropping patient_id, lesion_id, img_id, biopsed, diagnostic (:134-136)
    "smoke": TRI, "drink": TRI,
    "background_father": ["POMERANIA", "GERMANY", "BRAZIL", "NETHERLANDS", "ITALY", "POLAND", "PORTUGAL", "EMPTY", "UNK", "SPAIN", "AUSTRIA", "FRANCE", "CZECH"],
    "background_mother": ["POMERANIA", "GERMANY", "BRAZIL", "NETHERLANDS", "ITALY", "POLAND", "PORTUGAL", "EMPTY", "UNK", "SPAIN", "NORWAY"],
    "pesticide": TRI, "gender": ["FEMALE", "MALE", "EMPTY"], "skin_cancer_history": TRI, "cancer_history": TRI,
    "has_piped_water": TRI, "has_sewage_system": TRI, "fitspatrick": ["1.0", "2.0", "3.0", "4.0", "5.0", "6.0", "EMPTY"],
    "region": ["ARM", "NECK", "FACE", "HAND", "FOREARM", "CHEST", "NOSE", "THIGH", "SCALP", "EAR", "BACK", "FOOT", "ABDOMEN", "LIP"],
    "itch": ["True", "False", "UNK"], "grew": ["True", "False", "UNK"], "hurt": ["True", "False", "UNK"],
    "changed": ["True", "False", "UNK"], "bleed": ["True", "False", "UNK"], "elevation": ["True", "False", "UNK"],
}
NUMERICAL = ["age", "diameter_1", "diameter_2"]


def frame(rng, n, unknown=False):
    data = {}
    for col, vals in COLUMNS.items():
        pick = list(vals) + (["NEVER_SEEN", "ZZZ"] if unknown else [])
        data[col] = rng.choice(pick, size=n)
    data["age"] = rng.integers(6, 95, size=n).astype(np.float64)
    data["diameter_1"] = np.round(rng.gamma(2.0, 5.0, size=n), 1)
    data["diameter_2"] = np.round(rng.gamma(2.0, 4.0, size=n), 1)
    for col in ("diameter_1", "diameter_2"):
        data[col][rng.random(n) < 0.3] = np.nan               # missing measurements ("EMPTY" in the csv -> NaN -> -1, :147-152)
    return pd.DataFrame(data)


def encode_like_reference(df, ohe=None, scaler=None):
    cat_cols = [c for c in df.columns if c not in NUMERICAL]
    feats = df.copy()
    feats[cat_cols] = feats[cat_cols].astype(str)
    feats[NUMERICAL] = feats[NUMERICAL].apply(pd.to_numeric, errors="coerce").fillna(-1)
    if ohe is None:
        ohe = OneHotEncoder(sparse_output=False, handle_unknown="ignore")
        cat = ohe.fit_transform(feats[cat_cols])
        scaler = StandardScaler()
        num = scaler.fit_transform(feats[NUMERICAL])
    else:
        cat = ohe.transform(feats[cat_cols])
        num = scaler.transform(feats[NUMERICAL])
    return np.hstack((cat, num)), ohe, scaler, cat_cols


def main():
    rng = np.random.Generator(np.random.PCG64(2020))
    fit_df = frame(rng, 600)
    dense_fit, ohe, scaler, cat_cols = encode_like_reference(fit_df)
    test_df = frame(rng, 97, unknown=True)
    dense_test, _, _, _ = encode_like_reference(test_df, ohe, scaler)
    width = max(len(str(v)) for col in cat_cols for v in list(fit_df[col]) + list(test_df[col]))
    out = dict(
        fit_cat=fit_df[cat_cols].to_numpy().astype(f"<U{width}"), fit_num=fit_df[NUMERICAL].to_numpy(np.float64),
        test_cat=test_df[cat_cols].to_numpy().astype(f"<U{width}"), test_num=test_df[NUMERICAL].to_numpy(np.float64),
        dense_fit_head=dense_fit[:64].astype(np.float32), dense_test=dense_test.astype(np.float32),
        mean=scaler.mean_, scale=scaler.scale_, n_categories=np.array([len(c) for c in ohe.categories_], np.int32),
        categories_flat=np.concatenate([np.asarray(c, dtype=f"<U{width}") for c in ohe.categories_]),
    )
    np.savez_compressed(os.path.join(HERE, "metadata_pad20.npz"), **out)
    print("features:", dense_fit.shape[1], "categorical:", int(out["n_categories"].sum()), "file:", os.path.join(HERE, "metadata_pad20.npz"))


if __name__ == "__main__":
    main()
