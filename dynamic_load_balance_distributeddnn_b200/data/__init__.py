from .corpus import Corpus, Dictionary, SyntheticCorpus, batchify, get_batch, load_corpus
from .partition import DataPartitioner, Shard, global_permutation, split_token_stream
from .vision import SPECS, BatchStager, ImageDataset, load_image_dataset

__all__ = ["Corpus", "Dictionary", "SyntheticCorpus", "batchify", "get_batch", "load_corpus",
           "DataPartitioner", "Shard", "global_permutation", "split_token_stream",
           "SPECS", "BatchStager", "ImageDataset", "load_image_dataset"]
