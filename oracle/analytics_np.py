"""TEST INFRASTRUCTURE ONLY -- pandas restatement of the frame read-outs of the reference's
examples/lens_design.ipynb, the checker for pyrayt_b200/analytics.py.

Each function follows the notebook cell it cites, on the host frame the oracle / reference produced.
Nothing in the product path may import this module.
"""
import numpy as np
import pandas as pd


def _rows(results: pd.DataFrame, surface=None, generation=None) -> pd.DataFrame:
    if surface is not None:
        return results.loc[results["surface"] == surface]  # cell 11: results['surface'] == imager.get_id()
    if generation is not None:
        return results.loc[results["generation"] == generation]  # cell 12: generation == max(generation)
    return results


def focus(rows: pd.DataFrame) -> pd.Series:
    # cell 12: intercept = -x_tilt * y0 / y_tilt + x0
    with np.errstate(divide="ignore", invalid="ignore"):
        return -rows["x_tilt"] * rows["y0"] / rows["y_tilt"] + rows["x0"]


def spot_stats(results: pd.DataFrame, rays_per_group: int, n_groups: int, surface=None, generation=None,
               tilt_center=None) -> pd.DataFrame:
    """cells 11 / 19 / 38: per source_id, the spot of the selected rows (y1, z1) and their focus."""
    rows = _rows(results, surface, generation).copy()
    # RayTracer.calculate_source_ids (pyrayt/_pyrayt.py:316-327)
    rows["source_id"] = (rows["id"] / rays_per_group).astype(int)
    rows["focus"] = focus(rows)
    out = []
    for g in range(n_groups):
        sub = rows.loc[rows["source_id"] == g]
        y, z, t = np.asarray(sub["y1"]), np.asarray(sub["z1"]), np.asarray(sub["y_tilt"])
        f = np.asarray(sub["focus"])
        f = f[np.isfinite(f)]
        n = len(sub)
        if n == 0:
            out.append(dict(n=0, n_focus=0))
            continue
        ct = 0.0 if tilt_center is None else tilt_center
        out.append(dict(
            n=n, y_mean=np.mean(y), z_mean=np.mean(z), y_std=np.std(y), z_std=np.std(z),
            yz_cov=np.mean((y - np.mean(y)) * (z - np.mean(z))),
            rms_radius=np.sqrt(np.mean((y - np.mean(y)) ** 2 + (z - np.mean(z)) ** 2)),
            y_min=np.min(y), y_max=np.max(y), z_min=np.min(z), z_max=np.max(z),
            n_focus=len(f), focus_mean=np.mean(f) if len(f) else np.nan, focus_std=np.std(f) if len(f) else np.nan,
            y_tilt_mean=np.mean(t), y_tilt_std=np.std(t),
            sin_tilt_msd=np.mean(np.square(np.sin(t) - ct)),  # cell 20
        ))
    df = pd.DataFrame(out)
    df.index.name = "source_id"
    return df


def focus_table(results: pd.DataFrame, surface=None, generation=None) -> pd.DataFrame:
    """cells 12 and 15: focus of every selected ray against its launch radius and wavelength."""
    rows = _rows(results, surface, generation)
    gen0 = results.loc[results["generation"] == 0].set_index("id")["y0"]
    radius = np.asarray(gen0.reindex(np.asarray(rows["id"])))
    return pd.DataFrame({"id": np.asarray(rows["id"]), "radius": radius, "focus": np.asarray(focus(rows)),
                         "wavelength": np.asarray(rows["wavelength"])})
