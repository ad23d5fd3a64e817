"""Known-answer vectors transcribed from the reference's own unit tests for the predict path.

Data only (inputs + expected outputs) — the reference builds these models in code inside
its `#[cfg(test)]` modules; here each is a plain dict that tests encode with
`vpt_testlib.bincode_model.encode_model` and feed through `Model::read`.
Citations: /root/reference/vaporetto/src/<file>:<lines>  (SURVEY.md Appendix B numbering).

A scorer-only test of the reference (`CharScorerBoundary::new(...).add_scores` on a strip
pre-filled with `init`) is expressed as a whole model whose other scorer is absent
(`CharScorer::new`/`TypeScorer::new` return None for empty n-gram lists, char_scorer.rs:98,
type_scorer.rs:109) and whose bias is `init` (predictor.rs:520-524 fills the strip with bias).
"""

D, R, H, T, K, O = 1, 2, 3, 4, 5, 6  # CharacterType (sentence.rs:13-28)


def ty(*a):
    return bytes(a)


# --- predictor.rs:749-838 create_test_model -------------------------------------------------
PREDICTOR_TEST_MODEL = dict(
    char_ngrams=[("この人", [1, -2, 3, 4]), ("人だ", [-5, 6, 7, 8, 9])],
    type_ngrams=[(ty(H, H, K), [10, -11, 12, 13]), (ty(K, H), [-14, 15, 16, 17, -18])],
    dict=[("人", [19, 20], ""), ("地球", [21, -22, 23], "")],
    bias=5,
    char_window=3,
    type_window=3,
    tag_models=[
        dict(
            token="人",
            tags=[["名詞", "接尾辞"], ["ジン", "ヒト"]],
            char_ngrams=[("は地球人", [(0, [-32, 33, 34, -35])])],
            type_ngrams=[(ty(H, K, H), [(1, [36, -37, -38, 39])])],
            bias=[40, 41, 42, 43],
        ),
        dict(
            token="地球",
            tags=[["名詞"], ["マンホーム", "チキュー"]],
            char_ngrams=[("は地球人", [(1, [-44, 45])])],
            type_ngrams=[],
            bias=[46, 47],
        ),
    ],
)

# #1 predictor.rs:841-859 test_predict_boundaries ; #2 :863-903 test_predict_tags
PREDICT_BOUNDARIES = dict(
    model=PREDICTOR_TEST_MODEL,
    text="この人は地球人だ",
    scores=[-22, 54, 58, 43, -54, 68, 48],
    boundaries=[0, 1, 1, 1, 0, 1, 1],
    # tags(): n_tags=2, flattened [char][slot]
    tags=[None, None, None, None, "名詞", "ヒト", None, None, None, None, "名詞", "チキュー", "接尾辞", "ジン", None, None],
)

# #4 predictor.rs:678-747 PositionalWeight += : ((off_y, y), (off_x, x)) -> (off, w)
POSITIONAL_WEIGHT_ADD = [
    ((-2, [1, 2, 3, 4]), (4, [2, 4, 8]), (-2, [1, 2, 3, 4, 0, 0, 2, 4, 8])),
    ((-2, [1, 2, 3, 4]), (2, [2, 4, 8]), (-2, [1, 2, 3, 4, 2, 4, 8])),
    ((-2, [1, 2, 3, 4]), (0, [2, 4, 8]), (-2, [1, 2, 5, 8, 8])),
    ((-2, [1, 2, 3, 4]), (-1, [2, 4, 8]), (-2, [1, 4, 7, 12])),
    ((-2, [1, 2, 3, 4]), (-2, [2, 4, 8]), (-2, [3, 6, 11, 4])),
    ((-2, [1, 2, 3, 4]), (-4, [2, 4, 8]), (-4, [2, 4, 9, 2, 3, 4])),
    ((-2, [1, 2, 3, 4]), (-5, [2, 4, 8]), (-5, [2, 4, 8, 1, 2, 3, 4])),
    ((-2, [1, 2, 3, 4]), (-7, [2, 4, 8]), (-7, [2, 4, 8, 0, 0, 1, 2, 3, 4])),
]

# #5 char_scorer.rs:171-185 test_weight_merger : adds (ngram, offset, weights) -> merged list
CHAR_WEIGHT_MERGER = dict(
    adds=[("東京都", -3, [1, 2, 3, 4]), ("京都", -3, [2, 4, 6, 8, 10]), ("京都", -2, [3, 6, 9]), ("大阪", -2, [4, 8, 12])],
    merged=[("京都", -3, [2, 7, 12, 17, 10]), ("大阪", -2, [4, 8, 12]), ("東京都", -3, [3, 9, 15, 21, 10])],
)
# #10 type_scorer.rs:194-208
TYPE_WEIGHT_MERGER = dict(
    adds=[(b"eab", -3, [1, 2, 3, 4]), (b"ab", -3, [2, 4, 6, 8, 10]), (b"ab", -3, [3, 6, 9]), (b"cd", -2, [4, 8, 12])],
    merged=[(b"ab", -3, [5, 10, 15, 8, 10]), (b"cd", -2, [4, 8, 12]), (b"eab", -3, [6, 12, 18, 12, 10])],
)

_WARERA = "我らは全世界の国民"

# #6 char_scorer.rs:188-254 test_add_scores_1
CHAR_ADD_SCORES_1 = dict(
    model=dict(
        char_ngrams=[("我ら", [1, 2, 3, 4, 5]), ("全世界", [6, 7, 8, 9]), ("国民", [10, 11, 12, 13, 14]),
                     ("世界", [15, 16, 17, 18, 19]), ("界", [20, 21, 22, 23, 24, 25])],
        dict=[("全世界", [26, 27, 28, 29], ""), ("世界", [30, 31, 32], ""), ("世", [33, 34], "")],
        bias=1, char_window=3, type_window=3),
    text=_WARERA, scores=[4, 5, 73, 135, 141, 122, 55, 38])

# #7 char_scorer.rs:257-320 test_add_scores_2
CHAR_ADD_SCORES_2 = dict(
    model=dict(
        char_ngrams=[("我ら", [1, 2, 3]), ("全世界", [4, 5]), ("国民", [6, 7, 8]), ("世界", [9, 10, 11]),
                     ("界", [12, 13, 14, 15])],
        dict=[("全世界", [16, 17, 18, 19], ""), ("世界", [20, 21, 22], ""), ("世", [23, 24], "")],
        bias=2, char_window=2, type_window=2),
    text=_WARERA, scores=[4, 5, 18, 87, 93, 68, 23, 9])

# #8 char_scorer.rs:323-401 test_add_scores_3 (words longer than the window => Variable vectors)
CHAR_ADD_SCORES_3 = dict(
    model=dict(
        char_ngrams=[("我ら", [1, 2, 3, 4, 5]), ("全世界", [6, 7, 8, 9]), ("国民", [10, 11, 12, 13, 14]),
                     ("世界", [15, 16, 17, 18, 19]), ("界", [20, 21, 22, 23, 24, 25])],
        dict=[("全世界", [26, 27, 28, 29], ""), ("世界", [30, 31, 32], ""), ("世", [33, 34], ""),
              ("世界の国民", [35, 36, 37, 38, 39, 40], ""), ("は全世界", [41, 42, 43, 44, 45], "")],
        bias=3, char_window=3, type_window=3),
    text=_WARERA, scores=[6, 48, 117, 215, 223, 206, 95, 79])

# #9 char_scorer.rs:405-525 test_add_scores_with_tags (tag variant of the char scorer)
CHAR_ADD_SCORES_WITH_TAGS = dict(
    model=dict(
        char_ngrams=[("この人", [1, 2, 3, 4]), ("人だ", [5, 6, 7, 8, 9])],
        dict=[("人", [10, 11], ""), ("火星", [12, 13, 14], "")],
        bias=1, char_window=3, type_window=3,
        tag_models=[
            dict(token="t0", tags=[["a", "b", "c"]], bias=[0, 0, 0], type_ngrams=[],
                 char_ngrams=[("の人", [(0, [15, 16, 17]), (1, [18, 19, 20])]),
                              ("人は", [(1, [21, 22, 23]), (3, [24, 25, 26])]),
                              ("火星人", [(0, [27, 28, 29])])]),
            dict(token="t1", tags=[["a"]], bias=[], type_ngrams=[], char_ngrams=[]),
            dict(token="t2", tags=[["a", "b"]], bias=[0, 0], type_ngrams=[],
                 char_ngrams=[("人は", [(0, [27, 28]), (3, [29, 30])]), ("は火星人", [(3, [31, 32])])]),
        ]),
    text="この人は火星人だ",
    scores=[3, 14, 16, 13, 19, 31, 19],
    # (token_id, pos) -> tag scores starting from [1]*8
    tag_scores=[((0, 2), [37, 39, 41, 1, 1, 1, 1, 1]), ((0, 6), [28, 29, 30, 1, 1, 1, 1, 1]),
                ((2, 3), [59, 61, 1, 1, 1, 1, 1, 1])],
)

# #11 type_scorer.rs:211-259 test_add_scores (automaton variant: window 4 > CACHE_MAX_WINDOW_SIZE)
TYPE_ADD_SCORES = dict(
    model=dict(
        type_ngrams=[(ty(K, H), [1, 2, 3, 4, 5, 6, 7]), (ty(K, K, K), [8, 9, 10, 11, 12, 13]),
                     (ty(K, K), [14, 15, 16, 17, 18, 19, 20]), (ty(K), [21, 22, 23, 24, 25, 26, 27, 28])],
        bias=1, char_window=4, type_window=4),
    text=_WARERA, scores=[87, 135, 144, 174, 182, 192, 202, 148], type_variant=1)

# #12 type_scorer.rs:263-311 test_add_scores_cache_1
TYPE_ADD_SCORES_CACHE_1 = dict(
    model=dict(
        type_ngrams=[(ty(K, H), [1, 2, 3, 4, 5]), (ty(K, K, K), [6, 7, 8, 9]), (ty(K, K), [10, 11, 12, 13, 14]),
                     (ty(K), [15, 16, 17, 18, 19, 20])],
        bias=2, char_window=3, type_window=3),
    text=_WARERA, scores=[38, 66, 102, 84, 106, 139, 103, 74], type_variant=0)

# #13 type_scorer.rs:315-363 test_add_scores_cache_2
TYPE_ADD_SCORES_CACHE_2 = dict(
    model=dict(
        type_ngrams=[(ty(K, H), [1, 2, 3]), (ty(K, K, K), [4, 5]), (ty(K, K), [6, 7, 8]), (ty(K), [9, 10, 11, 12])],
        bias=3, char_window=2, type_window=2),
    text=_WARERA, scores=[16, 27, 28, 50, 57, 45, 43, 31], type_variant=0)

# #14 type_scorer.rs:367-473 test_add_scores_with_tags
TYPE_ADD_SCORES_WITH_TAGS = dict(
    model=dict(
        type_ngrams=[(ty(H, H, K), [1, 2, 3, 4]), (ty(K, H), [5, 6, 7, 8, 9])],
        bias=1, char_window=3, type_window=3,
        tag_models=[
            dict(token="t0", tags=[["a", "b", "c"]], bias=[0, 0, 0], char_ngrams=[],
                 type_ngrams=[(ty(H, K), [(0, [10, 11, 12]), (1, [13, 14, 15])]),
                              (ty(K, H), [(1, [16, 17, 18]), (3, [19, 20, 21])]),
                              (ty(K, K, K), [(0, [22, 23, 24])])]),
            dict(token="t1", tags=[["a"]], bias=[], char_ngrams=[], type_ngrams=[]),
            dict(token="t2", tags=[["a", "b"]], bias=[0, 0], char_ngrams=[],
                 type_ngrams=[(ty(K, H), [(0, [25, 26]), (3, [27, 28])]), (ty(H, K, K, K), [(3, [29, 30])])]),
        ]),
    text="この人は火星人だ",
    scores=[8, 10, 12, 9, 15, 7, 8],
    tag_scores=[((0, 2), [27, 29, 31, 1, 1, 1, 1, 1]), ((0, 6), [39, 41, 43, 1, 1, 1, 1, 1]),
                ((2, 3), [55, 57, 1, 1, 1, 1, 1, 1])],
    type_variant=1,
)

SCORE_CASES = {
    "predict_boundaries": PREDICT_BOUNDARIES,
    "char_add_scores_1": CHAR_ADD_SCORES_1,
    "char_add_scores_2": CHAR_ADD_SCORES_2,
    "char_add_scores_3": CHAR_ADD_SCORES_3,
    "type_add_scores": TYPE_ADD_SCORES,
    "type_add_scores_cache_1": TYPE_ADD_SCORES_CACHE_1,
    "type_add_scores_cache_2": TYPE_ADD_SCORES_CACHE_2,
}
# cases that need predict_tags=True to select the *Tag scorer variants
TAG_SCORE_CASES = {
    "predict_boundaries_tags": PREDICT_BOUNDARIES,
    "char_add_scores_with_tags": CHAR_ADD_SCORES_WITH_TAGS,
    "type_add_scores_with_tags": TYPE_ADD_SCORES_WITH_TAGS,
}

# #15 fixture resources/model.bin: lib.rs:17-41, predictor.rs:388-429, sentence.rs:1121-1137
MODEL_BIN_TOKENIZE = [
    ("まぁ社長は火星猫だ", False, "まぁ 社長 は 火星 猫 だ"),
    ("まぁ社長は火星猫だ", True, "まぁ/名詞/マー 社長/名詞/シャチョー は/助詞/ワ 火星/名詞/カセー 猫/名詞/ネコ だ/助動詞/ダ"),
    ("まぁ良いだろう", False, "まぁ 良い だろう"),
    ("まぁ良いだろう", True, "まぁ/副詞/マー 良い/形容詞/ヨイ だろう/助動詞/ダロー"),
]
# probe-computed in the survey session (not asserted by the reference itself)
MODEL_BIN_SCORES = {
    "まぁ社長は火星猫だ": [-20845, 18525, -22231, 26247, 41050, -21407, 32767, 26247],
    "まぁ良いだろう": [-20845, 22513, -24763, 15910, -20845, -21669],
}

# #17 vaporetto_tantivy/src/lib.rs:263-491 with test_model/model.zst (wsconst "" cases)
TANTIVY_TOKENIZE = [
    ("東京特許許可局", "東京 特許 許可 局"),
    ("123456円🤌🏿", "1 2 3 4 5 6 円 🤌 🏿"),
]

# #18 vaporetto_tantivy/src/lib.rs:160-199 token_stream = KyteaFullwidthFilter -> predict -> --wsconst post-filters ->
# tokens of the ORIGINAL text; tests lib.rs:255-260 (empty), :263-295, :298-365, :368-399 (wsconst "D") with
# test_model/model.zst (tests/golden/tantivy_model.bin).  (text, wsconst, tokens joined by ' ')
TANTIVY_PIPELINE = [
    ("東京特許許可局", "", "東京 特許 許可 局"),
    ("123456円🤌🏿", "", "1 2 3 4 5 6 円 🤌 🏿"),
    ("123456円🤌🏿", "D", "123456 円 🤌 🏿"),
]

# #19 sentence.rs:2695-2701 test_sentence_to_tokenized_string_escape: from_partial_annotation("火-星-猫|の| |生-態|\\-n")
# -> write_tokenized_text == "火星猫 の \\  生態 \\\\n" (Rust literals): ' ' and '\\' of a surface are escaped with '\\'.
TOKENIZED_ESCAPE = dict(
    text="火星猫の 生態\\n",
    boundaries=[0, 0, 1, 1, 1, 0, 1, 0],
    tokenized="火星猫 の \\  生態 \\\\n",
    # a model that produces exactly these boundaries: bias -1, +2 on the boundary after 猫, の, ' ', 態 (window 1:
    # a unigram's two weights land on the boundaries before and after the character)
    model=dict(char_ngrams=[("猫", [0, 2]), ("の", [0, 2]), (" ", [0, 2]), ("態", [0, 2])], type_ngrams=[], dict=[], bias=-1,
               char_window=1, type_window=1, tag_models=[]),
)
