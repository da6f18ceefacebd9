"""Golden vectors for the semantic-search post-filters, produced by EXECUTING the
reference's own functions.  The modules that hold them (lego_nlp_recommeder.py, hf_nlp_recommender.py) cannot be
imported here (langchain / sentence-transformers are not installed), so the two pure-Python methods are located in the
reference's source with ``ast``, compiled from the reference file where it lies and called unbound -- nothing is copied
into this repository; only inputs and outputs are stored.

    python oracle/make_golden_nlp.py        # writes tests/golden/nlp_filter_golden.json
"""
import ast
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/scripts"
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def extract_method(path, class_name, method):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == method:
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = {}
                    exec(compile(mod, path, "exec"), ns)
                    return ns[method]
    raise KeyError(f"{class_name}.{method} not found in {path}")


class Doc:
    def __init__(self, metadata):
        self.metadata = metadata
        self.page_content = ""


def main():
    import synth
    nlp_filter = extract_method(os.path.join(REF, "lego_nlp_recommeder.py"), "NLPRecommender", "_apply_filters")
    hf_filter = extract_method(os.path.join(REF, "hf_nlp_recommender.py"), "HuggingFaceNLPRecommender", "_apply_filters")
    rng = np.random.default_rng(12)
    themes = synth.themes_table()
    names = [themes[t] for t in sorted(themes)]
    n = 60
    metas = []
    for i in range(n):
        metas.append({"set_num": f"{i:05d}-1", "name": f"set {i}", "year": int(rng.integers(1995, 2026)),
                      "num_parts": int(np.exp(rng.uniform(0, np.log(5000)))), "theme": names[int(rng.integers(0, len(names)))],
                      "parent_theme": None, "num_colors": int(rng.integers(1, 30)), "num_minifigs": int(rng.integers(0, 8)),
                      "complexity": ["simple", "moderate", "complex"][int(rng.integers(0, 3))]})
    filter_sets = [
        {}, {"min_pieces": 100}, {"max_pieces": 300}, {"min_pieces": 50, "max_pieces": 800}, {"themes": ["city"]},
        {"themes": ["Star", "TECHNIC"]}, {"complexity": "moderate"}, {"min_age": 8}, {"min_age": 0},
        {"min_pieces": 0, "max_pieces": 0}, {"themes": []}, {"themes": ["space"], "min_age": 12, "max_pieces": 2000},
        {"complexity": "simple", "min_pieces": 20, "themes": ["Creator", "Friends"]}, {"themes": ["no such theme"]},
    ]
    nlp_cases = []
    for f in filter_sets:
        docs = [Doc(m) for m in metas]
        out = nlp_filter(None, docs, f)
        nlp_cases.append({"filters": f, "kept": [int(d.metadata["set_num"][:5]) for d in out]})
    hf_results = []
    for i, m in enumerate(metas):
        r = {"set_num": m["set_num"], "name": m["name"], "year": m["year"], "num_parts": m["num_parts"], "theme": m["theme"],
             "theme_id": i, "img_url": None, "relevance_score": 1.0 - i / n}
        if i % 3 == 0:
            r["min_age"], r["max_age"] = int(rng.integers(4, 16)), int(rng.integers(8, 99))
        if i % 7 == 0:
            r["num_parts"] = 0
        if i % 11 == 0:
            r["theme"] = None
        hf_results.append(r)
    hf_filter_sets = [{}, None, {"themes": [names[0], names[3]]}, {"min_pieces": 100}, {"max_pieces": 300}, {"min_age": 8},
                      {"max_age": 10}, {"themes": [names[1]], "min_pieces": 50, "max_age": 12}, {"themes": []}]
    hf_cases = []
    for f in hf_filter_sets:
        out = hf_filter(None, [dict(r) for r in hf_results], f)
        hf_cases.append({"filters": f, "kept": [int(r["set_num"][:5]) for r in out]})
    path = os.path.join(ROOT, "tests", "golden", "nlp_filter_golden.json")
    json.dump({"metas": metas, "nlp_cases": nlp_cases, "hf_results": hf_results, "hf_cases": hf_cases}, open(path, "w"), indent=0)
    print("wrote", path, len(nlp_cases), "nlp cases,", len(hf_cases), "hf cases")


if __name__ == "__main__":
    main()
