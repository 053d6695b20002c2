"""Feature-column API surface kept from deepctr_torch / CIRS (pure metadata, no compute).

Mirrors DeepCTR-Torch/deepctr_torch/inputs.py:20-120 (SparseFeat / VarLenSparseFeat / DenseFeat namedtuples,
build_input_features -> OrderedDict{name: (start, end)}, get_feature_names) and core/inputs.py:12-44
(SparseFeatP with padding_idx, get_dataset_columns per environment), core/user_model.py:538-557
(compute_input_dim).  Written from the interface description; semantics are identical so that the reference's
scripts can import these names unchanged.
"""
from collections import OrderedDict, namedtuple

DEFAULT_GROUP_NAME = "default_group"


class SparseFeat(namedtuple("SparseFeat", ["name", "vocabulary_size", "embedding_dim", "use_hash", "dtype",
                                           "embedding_name", "group_name"])):
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                group_name=DEFAULT_GROUP_NAME):
        if embedding_name is None:
            embedding_name = name
        if embedding_dim == "auto":
            embedding_dim = 6 * int(pow(vocabulary_size, 0.25))
        return super().__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype, embedding_name, group_name)

    def __hash__(self):
        return hash(self.name)


class SparseFeatP(SparseFeat):
    """SparseFeat + padding_idx (core/inputs.py:12-20)."""

    def __new__(cls, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                group_name=DEFAULT_GROUP_NAME, padding_idx=None):
        return super().__new__(cls, name, vocabulary_size, embedding_dim, use_hash, dtype, embedding_name, group_name)

    def __init__(self, name, vocabulary_size, embedding_dim=4, use_hash=False, dtype="int32", embedding_name=None,
                 group_name=DEFAULT_GROUP_NAME, padding_idx=None):
        self.padding_idx = padding_idx


class VarLenSparseFeat(namedtuple("VarLenSparseFeat", ["sparsefeat", "maxlen", "combiner", "length_name"])):
    __slots__ = ()

    def __new__(cls, sparsefeat, maxlen, combiner="mean", length_name=None):
        return super().__new__(cls, sparsefeat, maxlen, combiner, length_name)

    name = property(lambda self: self.sparsefeat.name)
    vocabulary_size = property(lambda self: self.sparsefeat.vocabulary_size)
    embedding_dim = property(lambda self: self.sparsefeat.embedding_dim)
    use_hash = property(lambda self: self.sparsefeat.use_hash)
    dtype = property(lambda self: self.sparsefeat.dtype)
    embedding_name = property(lambda self: self.sparsefeat.embedding_name)
    group_name = property(lambda self: self.sparsefeat.group_name)

    def __hash__(self):
        return hash(self.name)


class DenseFeat(namedtuple("DenseFeat", ["name", "dimension", "dtype"])):
    __slots__ = ()

    def __new__(cls, name, dimension=1, dtype="float32"):
        return super().__new__(cls, name, dimension, dtype)

    def __hash__(self):
        return hash(self.name)


def build_input_features(feature_columns):
    """OrderedDict {feature_name: (start, end)} column slices of the flat input matrix."""
    features, start = OrderedDict(), 0
    for feat in feature_columns:
        if feat.name in features:
            continue
        if isinstance(feat, SparseFeat):
            width = 1
        elif isinstance(feat, DenseFeat):
            width = feat.dimension
        elif isinstance(feat, VarLenSparseFeat):
            width = feat.maxlen
        else:
            raise TypeError(f"Invalid feature column type, got {type(feat)}")
        features[feat.name] = (start, start + width)
        start += width
        if isinstance(feat, VarLenSparseFeat) and feat.length_name is not None and feat.length_name not in features:
            features[feat.length_name] = (start, start + 1)
            start += 1
    return features


def get_feature_names(feature_columns):
    return list(build_input_features(feature_columns).keys())


def compute_input_dim(feature_columns, include_sparse=True, include_dense=True, feature_group=False):
    """core/user_model.py:538-557: summed embedding widths of sparse columns + widths of dense columns."""
    sparse = [f for f in feature_columns if isinstance(f, (SparseFeat, VarLenSparseFeat))]
    dense = [f for f in feature_columns if isinstance(f, DenseFeat)]
    dense_dim = sum(f.dimension for f in dense)
    sparse_dim = len(sparse) if feature_group else sum(f.embedding_dim for f in sparse)
    return (sparse_dim if include_sparse else 0) + (dense_dim if include_dense else 0)


def get_dataset_columns(dim_model, envname="VirtualTB-v0", env=None):
    """core/inputs.py:24-44.  Returns (user_columns, action_columns, feedback_columns, has_user_embedding,
    has_action_embedding, has_feedback_embedding)."""
    if envname == "VirtualTB-v0":
        return ([DenseFeat("feat_user", 88)], [DenseFeat("feat_item", 27)], [DenseFeat("feat_feedback", 1)],
                True, True, True)
    if envname == "KuaishouEnv-v0":
        n_user, n_item = env.mat.shape[0], env.mat.shape[1]
        return ([SparseFeatP("feat_user", n_user, embedding_dim=dim_model)],
                [SparseFeatP("feat_item", n_item, embedding_dim=dim_model)],
                [DenseFeat("feat_feedback", 1)], False, False, True)
    return [], [], [], None, None, None
