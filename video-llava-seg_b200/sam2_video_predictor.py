"""`SAM2VideoPredictor` with the public surface of sam2/sam2_video_predictor.py: init_state,
add_new_points_or_box, add_new_mask, propagate_in_video, clear_all_prompts_in_frame, reset_state --
same arguments, same `(frame_idx, obj_ids, video_res_masks)` results, same RuntimeError / ValueError
behaviour -- on top of the libvls_b200 tracking core.

`init_state(video_path=...)` additionally accepts (a) a [T,3,H,W] tensor of normalised frames and
(b) any object with `num_frames`, `video_height`, `video_width` and `frame_features(t, device)`
returning the dict SAM2Base.forward_image would produce (precomputed backbone features; this is how
bench.py keeps the out-of-scope image encoder off the timed hot path).
"""
import warnings
from collections import OrderedDict

import torch

from . import ops
from .graphed import FrameGraph, SteadyStateGraph, baked_settings
from .modeling.sam2_base import NO_OBJ_SCORE, SAM2Base
from .utils.misc import fill_holes_in_mask_scores, load_video_frames


def _concat_points(old, pts, labels):
    if old is None:
        return {"point_coords": pts, "point_labels": labels}
    return {"point_coords": torch.cat([old["point_coords"], pts], dim=1),
            "point_labels": torch.cat([old["point_labels"], labels], dim=1)}


def _empty_frame_store():
    return {"cond_frame_outputs": {}, "non_cond_frame_outputs": {}}


class _GraphOwner:
    """Lifetime token of a session (plain dicts cannot be weakly referenced)."""


class SAM2VideoPredictor(SAM2Base):
    def __init__(self, fill_hole_area=0, non_overlap_masks=False, clear_non_cond_mem_around_input=False,
                 clear_non_cond_mem_for_multi_obj=False, add_all_frames_to_correct_as_cond=False, **kwargs):
        super().__init__(**kwargs)
        self.fill_hole_area = fill_hole_area
        self.non_overlap_masks = non_overlap_masks
        self.clear_non_cond_mem_around_input = clear_non_cond_mem_around_input
        self.clear_non_cond_mem_for_multi_obj = clear_non_cond_mem_for_multi_obj
        self.add_all_frames_to_correct_as_cond = add_all_frames_to_correct_as_cond
        # replay the steady-state frame as one CUDA graph (graphed.py); set to False to force the eager path
        self.use_cuda_graph = True
        # what propagate_in_video / add_new_* yield as `video_res_masks`:
        #   "logits" f32 [B,1,Hv,Wv] scores, as the reference (sam2_video_predictor.py:404-424)
        #   "binary" uint8 0/1 [B,1,Hv,Wv] = (scores > 0), "bits" the same bit-packed [B,1,Hv,ceil(Wv/8)] (np.packbits
        #   order): the fused resize+threshold output stage; the f32 video-resolution logits are
        #   never written.  Not available with non_overlap_masks (which needs the scores of every object).
        self.output_mode = "logits"

    # ------------------------------------------------------------------ session
    @torch.inference_mode()
    def init_state(self, video_path, offload_video_to_cpu=False, offload_state_to_cpu=False,
                   async_loading_frames=False):
        dev = self.device
        st = {}
        if hasattr(video_path, "frame_features"):
            st["feature_source"], st["images"] = video_path, None
            st["num_frames"] = video_path.num_frames
            h, w = video_path.video_height, video_path.video_width
        else:
            images, h, w = load_video_frames(video_path=video_path, image_size=self.image_size,
                                             offload_video_to_cpu=offload_video_to_cpu,
                                             async_loading_frames=async_loading_frames, compute_device=dev)
            st["feature_source"], st["images"], st["num_frames"] = None, images, len(images)
        st["offload_video_to_cpu"], st["offload_state_to_cpu"] = offload_video_to_cpu, offload_state_to_cpu
        st["video_height"], st["video_width"], st["device"] = h, w, dev
        st["storage_device"] = torch.device("cpu") if offload_state_to_cpu else dev
        st["point_inputs_per_obj"], st["mask_inputs_per_obj"] = {}, {}
        st["cached_features"], st["constants"] = {}, {}
        st["obj_id_to_idx"], st["obj_idx_to_id"], st["obj_ids"] = OrderedDict(), OrderedDict(), []
        st["output_dict"] = _empty_frame_store()
        st["output_dict_per_obj"], st["temp_output_dict_per_obj"] = {}, {}
        st["consolidated_frame_inds"] = {"cond_frame_outputs": set(), "non_cond_frame_outputs": set()}
        st["tracking_has_started"], st["frames_already_tracked"] = False, {}
        st["steady_graph"] = None
        self._get_image_feature(st, frame_idx=0, batch_size=1)  # warm up / cache frame 0, as the reference does
        return st

    def _obj_id_to_idx(self, st, obj_id):
        idx = st["obj_id_to_idx"].get(obj_id, None)
        if idx is not None:
            return idx
        if st["tracking_has_started"]:
            raise RuntimeError(f"Cannot add new object id {obj_id} after tracking starts. "
                               f"All existing object ids: {st['obj_ids']}. "
                               f"Please call 'reset_state' to restart from scratch.")
        idx = len(st["obj_id_to_idx"])
        st["obj_id_to_idx"][obj_id], st["obj_idx_to_id"][idx] = idx, obj_id
        st["obj_ids"] = list(st["obj_id_to_idx"])
        st["point_inputs_per_obj"][idx], st["mask_inputs_per_obj"][idx] = {}, {}
        st["output_dict_per_obj"][idx], st["temp_output_dict_per_obj"][idx] = _empty_frame_store(), _empty_frame_store()
        return idx

    def _obj_idx_to_id(self, st, obj_idx):
        return st["obj_idx_to_id"][obj_idx]

    def _get_obj_num(self, st):
        return len(st["obj_idx_to_id"])

    # ------------------------------------------------------------------ prompts
    def _prompt_frame(self, st, frame_idx, obj_idx, point_inputs, mask_inputs):
        """Shared tail of add_new_points_or_box / add_new_mask (sam2_video_predictor.py:250-314,355-402)."""
        self._drop_graph(st)  # new prompts change the memory bank
        is_init = frame_idx not in st["frames_already_tracked"]
        reverse = False if is_init else st["frames_already_tracked"][frame_idx]["reverse"]
        obj_out, obj_tmp = st["output_dict_per_obj"][obj_idx], st["temp_output_dict_per_obj"][obj_idx]
        is_cond = is_init or self.add_all_frames_to_correct_as_cond
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        prev_logits = None
        if point_inputs is not None and "prompt_embedding" not in point_inputs:
            prev = obj_tmp[key].get(frame_idx) or obj_out["cond_frame_outputs"].get(frame_idx) \
                or obj_out["non_cond_frame_outputs"].get(frame_idx)
            if prev is not None and prev["pred_masks"] is not None:
                prev_logits = torch.clamp(prev["pred_masks"].to(st["device"], non_blocking=True), -32.0, 32.0)
        cur, _ = self._run_single_frame_inference(
            inference_state=st, output_dict=obj_out, frame_idx=frame_idx, batch_size=1, is_init_cond_frame=is_init,
            point_inputs=point_inputs, mask_inputs=mask_inputs, reverse=reverse, run_mem_encoder=False,
            prev_sam_mask_logits=prev_logits)
        obj_tmp[key][frame_idx] = cur
        out = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond, run_mem_encoder=False,
                                                       consolidate_at_video_res=True)
        _, video_res = self._get_orig_video_res_output(st, out["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res

    @torch.inference_mode()
    def add_new_points_or_box(self, inference_state, frame_idx, obj_id, points=None, labels=None,
                              clear_old_points=True, normalize_coords=True, box=None):
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        if (points is not None) != (labels is not None):
            raise ValueError("points and labels must be provided together")
        if points is None and box is None:
            raise ValueError("at least one of points or box must be provided as input")
        points = torch.zeros(0, 2, dtype=torch.float32) if points is None else torch.as_tensor(points, dtype=torch.float32)
        labels = torch.zeros(0, dtype=torch.int32) if labels is None else torch.as_tensor(labels, dtype=torch.int32)
        if points.dim() == 2:
            points = points.unsqueeze(0)
        if labels.dim() == 1:
            labels = labels.unsqueeze(0)
        if box is not None:
            if not clear_old_points:
                raise ValueError("cannot add box without clearing old points, since box prompt must be provided "
                                 "before any point prompt (please use clear_old_points=True instead)")
            if st["tracking_has_started"]:
                warnings.warn("You are adding a box after tracking starts. SAM 2 may not always be able to incorporate "
                              "a box prompt for *refinement*; call 'reset_state' to restart from scratch if this is an "
                              "initial prompt.", category=UserWarning, stacklevel=2)
            box = torch.as_tensor(box, dtype=torch.float32, device=points.device).reshape(1, 2, 2)
            points = torch.cat([box, points], dim=1)
            labels = torch.cat([torch.tensor([[2, 3]], dtype=torch.int32, device=labels.device), labels], dim=1)
        if normalize_coords:
            points = points / torch.tensor([st["video_width"], st["video_height"]]).to(points.device)
        points = (points * self.image_size).to(st["device"])
        labels = labels.to(st["device"])
        per_frame = st["point_inputs_per_obj"][obj_idx]
        point_inputs = _concat_points(None if clear_old_points else per_frame.get(frame_idx, None), points, labels)
        per_frame[frame_idx] = point_inputs
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        return self._prompt_frame(st, frame_idx, obj_idx, point_inputs, None)

    @torch.inference_mode()
    def add_new_prompt_embedding(self, inference_state, frame_idx, obj_id, sparse_embedding):
        """Prompt an object with a sparse prompt EMBEDDING instead of clicks: `sparse_embedding` [Ns,256] or
        [1,Ns,256], e.g. the projected `[SEG]` hidden state of the LLM (llava/model/seg_head/sam2.py:75-88).  The
        frame becomes a conditioning frame exactly as with add_new_points_or_box; propagate_in_video then tracks
        the object (BASELINE config 4: "[SEG] prompt -> decoder + propagation")."""
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        e = torch.as_tensor(sparse_embedding, dtype=torch.float32)
        if e.dim() == 2:
            e = e.unsqueeze(0)
        if e.dim() != 3 or e.shape[0] != 1 or e.shape[2] != self.hidden_dim:
            raise ValueError(f"sparse_embedding must be [Ns,{self.hidden_dim}] or [1,Ns,{self.hidden_dim}]")
        point_inputs = {"prompt_embedding": e.to(st["device"])}
        st["point_inputs_per_obj"][obj_idx][frame_idx] = point_inputs
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        return self._prompt_frame(st, frame_idx, obj_idx, point_inputs, None)

    def add_new_points(self, *args, **kwargs):
        return self.add_new_points_or_box(*args, **kwargs)

    @torch.inference_mode()
    def add_new_mask(self, inference_state, frame_idx, obj_id, mask):
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        mask = torch.as_tensor(mask, dtype=torch.bool) if not isinstance(mask, torch.Tensor) else mask
        assert mask.dim() == 2
        m = mask[None, None].float().to(st["device"])
        if tuple(mask.shape) != (self.image_size, self.image_size):
            m = torch.nn.functional.interpolate(m, size=(self.image_size, self.image_size), align_corners=False,
                                                mode="bilinear", antialias=True)
            m = (m >= 0.5).float()
        st["mask_inputs_per_obj"][obj_idx][frame_idx] = m
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        return self._prompt_frame(st, frame_idx, obj_idx, None, m)

    # ------------------------------------------------------------------ outputs
    def _get_orig_video_res_output(self, st, any_res_masks):
        """Resize scores to the video resolution (sam2_video_predictor.py:404-424) with the resize kernel."""
        any_res_masks = any_res_masks.to(st["device"], non_blocking=True)
        hw = (st["video_height"], st["video_width"])
        return any_res_masks, self._video_res_output(any_res_masks, hw)

    def _video_res_output(self, masks, hw):
        if self.output_mode != "logits":
            if self.output_mode not in ("binary", "bits"):
                raise ValueError(f"output_mode must be 'logits', 'binary' or 'bits', got {self.output_mode!r}")
            if self.non_overlap_masks:
                raise ValueError("output_mode='binary'/'bits' cannot be combined with non_overlap_masks")
            return ops.resize_binarize(masks, hw, 0.0, packed=self.output_mode == "bits")
        video_res = masks if tuple(masks.shape[-2:]) == hw else ops.resize_bilinear(masks, hw)
        if self.non_overlap_masks:
            video_res = self._apply_non_overlapping_constraints(video_res)
        return video_res

    def _consolidate_temp_output_across_obj(self, st, frame_idx, is_cond, run_mem_encoder,
                                            consolidate_at_video_res=False):
        """Merge per-object temporary outputs on a frame (sam2_video_predictor.py:426-554)."""
        B = self._get_obj_num(st)
        key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
        if consolidate_at_video_res:
            assert not run_mem_encoder, "memory encoder cannot run at video resolution"
            H, W, mask_key = st["video_height"], st["video_width"], "pred_masks_video_res"
        else:
            H = W = self.image_size // 4
            mask_key = "pred_masks"
        out = {
            "maskmem_features": None, "maskmem_pos_enc": None, "maskmem_rows": None,
            mask_key: torch.full((B, 1, H, W), NO_OBJ_SCORE, dtype=torch.float32, device=st["storage_device"]),
            "obj_ptr": torch.full((B, self.hidden_dim), NO_OBJ_SCORE, dtype=torch.float32, device=st["device"]),
            "object_score_logits": torch.full((B, 1), 10.0, dtype=torch.float32, device=st["device"]),
        }
        empty_ptr = None
        for i in range(B):
            tmp, per = st["temp_output_dict_per_obj"][i], st["output_dict_per_obj"][i]
            o = tmp[key].get(frame_idx) or per["cond_frame_outputs"].get(frame_idx) \
                or per["non_cond_frame_outputs"].get(frame_idx)
            if o is None:
                if run_mem_encoder:
                    if empty_ptr is None:
                        empty_ptr = self._get_empty_mask_ptr(st, frame_idx)
                    out["obj_ptr"][i:i + 1] = empty_ptr
                continue
            m = o["pred_masks"]
            if tuple(m.shape[-2:]) != (H, W):
                m = ops.resize_bilinear(m.to(st["device"]), (H, W)).to(out[mask_key].device)
            out[mask_key][i:i + 1] = m
            out["obj_ptr"][i:i + 1] = o["obj_ptr"]
            out["object_score_logits"][i:i + 1] = o["object_score_logits"]
        if run_mem_encoder:
            low = out["pred_masks"].to(st["device"], non_blocking=True)
            if self.non_overlap_masks_for_mem_enc:
                hi = self._apply_non_overlapping_constraints(ops.resize_bilinear(low, (self.image_size,) * 2))
                feats, rows, pos = self._run_memory_encoder(st, frame_idx, B, hi, out["object_score_logits"], True)
            else:
                # up-sampling of the consolidated low-res logits (:535-540) is fused into the memory encoder
                feats, rows, pos = self._run_memory_encoder(st, frame_idx, B, None, out["object_score_logits"], True,
                                                            low_res_masks=low)
            out["maskmem_features"], out["maskmem_rows"], out["maskmem_pos_enc"] = feats, rows, pos
        return out

    def _get_empty_mask_ptr(self, st, frame_idx):
        """Dummy object pointer from an empty mask (sam2_video_predictor.py:556-590)."""
        mask = torch.zeros((1, 1, self.image_size, self.image_size), dtype=torch.float32, device=st["device"])
        _, _, feats, pos, sizes = self._get_image_feature(st, frame_idx, 1)
        cur = self.track_step(frame_idx=frame_idx, is_init_cond_frame=True, current_vision_feats=feats,
                              current_vision_pos_embeds=pos, feat_sizes=sizes, point_inputs=None, mask_inputs=mask,
                              output_dict={}, num_frames=st["num_frames"], track_in_reverse=False,
                              run_mem_encoder=False, prev_sam_mask_logits=None)
        return cur["obj_ptr"]

    # ------------------------------------------------------------------ propagation
    @torch.inference_mode()
    def propagate_in_video_preflight(self, st):
        """Consolidate prompt-frame outputs and encode their memories (sam2_video_predictor.py:592-660)."""
        st["tracking_has_started"] = True
        B = self._get_obj_num(st)
        tmp_all, out_all, done = st["temp_output_dict_per_obj"], st["output_dict"], st["consolidated_frame_inds"]
        for is_cond in (False, True):
            key = "cond_frame_outputs" if is_cond else "non_cond_frame_outputs"
            frames = set()
            for tmp in tmp_all.values():
                frames.update(tmp[key].keys())
            done[key].update(frames)
            for f in frames:
                merged = self._consolidate_temp_output_across_obj(st, f, is_cond=is_cond, run_mem_encoder=True)
                out_all[key][f] = merged
                self._add_output_per_object(st, f, merged, key)
                if self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1):
                    self._clear_non_cond_mem_around_input(st, f)
            for tmp in tmp_all.values():
                tmp[key].clear()
        for f in out_all["cond_frame_outputs"]:
            out_all["non_cond_frame_outputs"].pop(f, None)
        for per in st["output_dict_per_obj"].values():
            for f in per["cond_frame_outputs"]:
                per["non_cond_frame_outputs"].pop(f, None)
        for f in done["cond_frame_outputs"]:
            assert f in out_all["cond_frame_outputs"]
            done["non_cond_frame_outputs"].discard(f)
        prompted = set()
        for d in list(st["point_inputs_per_obj"].values()) + list(st["mask_inputs_per_obj"].values()):
            prompted.update(d.keys())
        assert (done["cond_frame_outputs"] | done["non_cond_frame_outputs"]) == prompted

    @torch.inference_mode()
    def propagate_in_video(self, inference_state, start_frame_idx=None, max_frame_num_to_track=None, reverse=False):
        """Generator of (frame_idx, obj_ids, video_res_masks) (sam2_video_predictor.py:662-745)."""
        st = inference_state
        self.propagate_in_video_preflight(st)
        out_all, done, obj_ids = st["output_dict"], st["consolidated_frame_inds"], st["obj_ids"]
        T, B = st["num_frames"], self._get_obj_num(st)
        if len(out_all["cond_frame_outputs"]) == 0:
            raise RuntimeError("No points are provided; please add points first")
        clear_mem = self.clear_non_cond_mem_around_input and (self.clear_non_cond_mem_for_multi_obj or B <= 1)
        if start_frame_idx is None:
            start_frame_idx = min(out_all["cond_frame_outputs"])
        if max_frame_num_to_track is None:
            max_frame_num_to_track = T
        if reverse:
            end = max(start_frame_idx - max_frame_num_to_track, 0)
            order = range(start_frame_idx, end - 1, -1) if start_frame_idx > 0 else []
        else:
            end = min(start_frame_idx + max_frame_num_to_track, T - 1)
            order = range(start_frame_idx, end + 1)
        for f in order:
            if f in done["cond_frame_outputs"]:
                key = "cond_frame_outputs"
                cur = out_all[key][f]
                pred = cur["pred_masks"]
                if clear_mem:
                    self._clear_non_cond_mem_around_input(st, f)
            elif f in done["non_cond_frame_outputs"]:
                key = "non_cond_frame_outputs"
                cur = out_all[key][f]
                pred = cur["pred_masks"]
            else:
                key = "non_cond_frame_outputs"
                g = st.get("steady_graph")
                if g is not None and not g.owned_by(st.get("graph_owner")):
                    g = st["steady_graph"] = None     # handed to another clip since this session last used it
                if g is not None and (g.next_frame != f or g.B != B or not g.valid()):
                    self._drop_graph(st)
                    g = None
                if g is None and SteadyStateGraph.eligible(self, st, f, B, reverse):
                    g = st["steady_graph"] = self._acquire_graph(st, f, B)
                fg = None if g is not None else self._frame_graph_step(st, f, B, reverse)
                if g is not None:   # full memory bank: one CUDA-graph replay per frame
                    cur, video_res = g.run(st, f)
                    pred = None
                elif fg is not None:   # growing bank of a shape seen before (ramp frames): replay too
                    cur, video_res = fg
                    pred = None
                else:
                    cur, pred = self._run_single_frame_inference(
                        inference_state=st, output_dict=out_all, frame_idx=f, batch_size=B, is_init_cond_frame=False,
                        point_inputs=None, mask_inputs=None, reverse=reverse, run_mem_encoder=True)
                out_all[key][f] = cur
            self._add_output_per_object(st, f, cur, key)
            st["frames_already_tracked"][f] = {"reverse": reverse}
            if pred is not None:
                _, video_res = self._get_orig_video_res_output(st, pred)
            yield f, obj_ids, video_res

    def _frame_graph_step(self, st, frame_idx, batch_size, reverse):
        """Track `frame_idx` through a FrameGraph if its memory-bank shape has been seen before (so that a capture,
        ~20+ ms, is only paid for shapes that recur: the ramp frames of the second and later clips).  Returns
        (compact state entry, video-resolution output) or None when the eager path must be used."""
        if not self.use_cuda_graph or st["offload_state_to_cpu"] or self.non_overlap_masks \
                or self.non_overlap_masks_for_mem_enc or self.num_maskmem == 0 or self.training:
            return None
        out_all = st["output_dict"]
        if not out_all["cond_frame_outputs"]:
            return None
        dev = st["device"]
        _, bo, _, _, _ = self._get_image_feature(st, frame_idx, 1)
        fpn, pe = bo["backbone_fpn"], bo["vision_pos_enc"]
        mem_parts, pos_parts, n_ptr_tokens = self._gather_bank(frame_idx, out_all, st["num_frames"], reverse, batch_size, dev)
        if any(p.shape[0] != batch_size for p in mem_parts):
            return None
        Nk = sum(p.shape[1] for p in mem_parts)
        hw = (st["video_height"], st["video_width"])
        key = (batch_size, Nk, n_ptr_tokens, hw, baked_settings(self), tuple(fpn[-3].shape), tuple(fpn[-2].shape),
               tuple(fpn[-1].shape), fpn[-1].dtype)
        cache = self.__dict__.setdefault("_frame_graphs", {})
        seen = self.__dict__.setdefault("_frame_shapes_seen", {})
        seen[key] = seen.get(key, 0) + 1
        g = cache.get(key)
        if g is not None and not g.valid():      # weights re-packed / workspaces re-allocated since the capture
            del cache[key]
            g = None
        if g is None:
            if seen[key] < 2 or len(cache) >= 64:
                return None
            g = cache[key] = FrameGraph(self, batch_size, Nk, n_ptr_tokens, hw, [fpn[-3], fpn[-2], fpn[-1], pe[-1]])
        pred, obj_ptr, obj_logits, nchw, rows, video = g.run(mem_parts, pos_parts, fpn, pe)
        compact = {
            "maskmem_features": nchw, "maskmem_rows": rows,
            "maskmem_pos_enc": self._get_maskmem_pos_enc(st, {"maskmem_pos_enc": [
                self._constants()["maskmem_pos"].expand(batch_size, -1, -1, -1)]}),
            "pred_masks": pred, "obj_ptr": obj_ptr, "object_score_logits": obj_logits,
        }
        return compact, video

    @staticmethod
    def _drop_graph(st):
        """The session stops using its captured graph (new prompts, object removed, ...): the graph becomes re-usable."""
        g = st.get("steady_graph")
        if g is not None:
            g.release(st.get("graph_owner"))     # no-op when the graph already serves another session
        st["steady_graph"] = None

    def _acquire_graph(self, st, frame_idx, batch_size):
        """A captured steady-state graph for this session: an idle one of the same shape is re-used (its static bank is
        refilled in place), otherwise a new one is built and remembered.  Sessions own a token whose lifetime tells the
        cache when a graph is free again."""
        if "graph_owner" not in st:
            st["graph_owner"] = _GraphOwner()
        cache = self.__dict__.setdefault("_graph_cache", {})
        key = SteadyStateGraph.key_for(self, st, frame_idx, batch_size)
        for g in cache.get(key, []):
            if g.idle() and g.valid():
                g.rebind(st, frame_idx, st["graph_owner"])
                return g
        g = SteadyStateGraph(self, st, frame_idx, batch_size, owner=st["graph_owner"])
        graphs = cache.setdefault(key, [])
        graphs[:] = [x for x in graphs if x.valid()][-3:]      # bound the cache: at most 4 graphs per shape
        graphs.append(g)
        return g

    def _add_output_per_object(self, st, frame_idx, cur, key):
        """Per-object views sharing storage with the batched output (sam2_video_predictor.py:747-774)."""
        feats, pos, rows = cur["maskmem_features"], cur["maskmem_pos_enc"], cur.get("maskmem_rows")
        for i, per in st["output_dict_per_obj"].items():
            s = slice(i, i + 1)
            per[key][frame_idx] = {
                "maskmem_features": None if feats is None else feats[s],
                "maskmem_rows": None if rows is None else rows[s],
                "maskmem_pos_enc": None if pos is None else [x[s] for x in pos],
                "pred_masks": cur["pred_masks"][s], "obj_ptr": cur["obj_ptr"][s],
                "object_score_logits": cur["object_score_logits"][s],
            }

    @torch.inference_mode()
    def clear_all_prompts_in_frame(self, inference_state, frame_idx, obj_id, need_output=True):
        st = inference_state
        obj_idx = self._obj_id_to_idx(st, obj_id)
        st["point_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        st["mask_inputs_per_obj"][obj_idx].pop(frame_idx, None)
        tmp_all = st["temp_output_dict_per_obj"]
        tmp_all[obj_idx]["cond_frame_outputs"].pop(frame_idx, None)
        tmp_all[obj_idx]["non_cond_frame_outputs"].pop(frame_idx, None)
        B = self._get_obj_num(st)
        has_input = any(frame_idx in st["point_inputs_per_obj"][i] or frame_idx in st["mask_inputs_per_obj"][i]
                        for i in range(B))
        if not has_input:
            out_all, done = st["output_dict"], st["consolidated_frame_inds"]
            done["cond_frame_outputs"].discard(frame_idx)
            done["non_cond_frame_outputs"].discard(frame_idx)
            o = out_all["cond_frame_outputs"].pop(frame_idx, None)
            if o is not None:
                out_all["non_cond_frame_outputs"][frame_idx] = o
                st["frames_already_tracked"].pop(frame_idx, None)
            for i in range(B):
                per = st["output_dict_per_obj"][i]
                o = per["cond_frame_outputs"].pop(frame_idx, None)
                if o is not None:
                    per["non_cond_frame_outputs"][frame_idx] = o
            if len(out_all["cond_frame_outputs"]) == 0:
                self._reset_tracking_results(st)
        if not need_output:
            return
        is_cond = any(frame_idx in t["cond_frame_outputs"] for t in tmp_all.values())
        out = self._consolidate_temp_output_across_obj(st, frame_idx, is_cond=is_cond, run_mem_encoder=False,
                                                       consolidate_at_video_res=True)
        _, video_res = self._get_orig_video_res_output(st, out["pred_masks_video_res"])
        return frame_idx, st["obj_ids"], video_res

    @classmethod
    def from_pretrained(cls, model_id, **kwargs):
        """sam2_video_predictor.py:33-43 downloads a checkpoint from the Hugging Face hub; this package is offline
        by design: build with build_sam.build_sam2_video_predictor(config, state_dict_or_checkpoint_path)."""
        raise RuntimeError(f"from_pretrained({model_id!r}) needs network access; load a local checkpoint with "
                           "video_llava_seg_b200.build_sam.build_sam2_video_predictor instead")

    @torch.inference_mode()
    def remove_object(self, inference_state, obj_id, strict=False, need_output=True):
        """Drop one tracked object (sam2_video_predictor.py:1042-1153): its prompts are cleared (which may demote
        conditioning frames), the object indices are re-packed, and every stored frame keeps only the remaining
        objects' rows.  Returns (remaining obj_ids, [(frame_idx, video_res_masks)] for the frames it had inputs on)."""
        st = inference_state
        rm = st["obj_id_to_idx"].get(obj_id)
        updated = []
        if rm is None:
            if strict:
                raise RuntimeError(f"Cannot remove object id {obj_id} as it doesn't exist. "
                                   f"All existing object ids: {st['obj_ids']}.")
            return st["obj_ids"], updated
        if len(st["obj_id_to_idx"]) == 1:
            self.reset_state(st)
            return st["obj_ids"], updated
        self._drop_graph(st)                          # the captured graph is specialised on the object count
        input_frames = set(st["point_inputs_per_obj"][rm]) | set(st["mask_inputs_per_obj"][rm])
        for f in input_frames:
            self.clear_all_prompts_in_frame(st, f, obj_id, need_output=False)
        old_ids = list(st["obj_ids"])
        keep = [i for i in range(len(old_ids)) if i != rm]
        new_ids = [old_ids[i] for i in keep]
        remap = {old: new for new, old in enumerate(keep)}
        st["obj_id_to_idx"] = {oid: i for i, oid in enumerate(new_ids)}
        st["obj_idx_to_id"] = {i: oid for i, oid in enumerate(new_ids)}
        st["obj_ids"] = new_ids
        for name in ("point_inputs_per_obj", "mask_inputs_per_obj", "output_dict_per_obj", "temp_output_dict_per_obj"):
            box = st[name]
            moved = {remap[k]: v for k, v in box.items() if k in remap}
            box.clear()
            box.update(moved)
        for key in ("cond_frame_outputs", "non_cond_frame_outputs"):
            for f, out in st["output_dict"][key].items():
                for field in ("maskmem_features", "maskmem_rows", "pred_masks", "obj_ptr", "object_score_logits"):
                    if out.get(field) is not None:
                        out[field] = out[field][keep]
                if out.get("maskmem_pos_enc") is not None:
                    out["maskmem_pos_enc"] = self._get_maskmem_pos_enc(st, {"maskmem_pos_enc": [
                        x[keep] if x.shape[0] == len(old_ids) else x for x in out["maskmem_pos_enc"]]})
                self._add_output_per_object(st, f, out, key)
        if need_output:
            tmp_all = st["temp_output_dict_per_obj"]
            for f in sorted(input_frames):
                is_cond = any(f in t["cond_frame_outputs"] for t in tmp_all.values())
                out = self._consolidate_temp_output_across_obj(st, f, is_cond=is_cond, run_mem_encoder=False,
                                                               consolidate_at_video_res=True)
                _, video_res = self._get_orig_video_res_output(st, out["pred_masks_video_res"])
                updated.append((f, video_res))
        return st["obj_ids"], updated

    @torch.inference_mode()
    def reset_state(self, inference_state):
        st = inference_state
        self._reset_tracking_results(st)
        for k in ("obj_id_to_idx", "obj_idx_to_id", "obj_ids", "point_inputs_per_obj", "mask_inputs_per_obj",
                  "output_dict_per_obj", "temp_output_dict_per_obj"):
            st[k].clear()

    def _reset_tracking_results(self, st):
        self._drop_graph(st)
        for k in ("point_inputs_per_obj", "mask_inputs_per_obj"):
            for v in st[k].values():
                v.clear()
        for k in ("output_dict_per_obj", "temp_output_dict_per_obj"):
            for v in st[k].values():
                v["cond_frame_outputs"].clear()
                v["non_cond_frame_outputs"].clear()
        for k in ("cond_frame_outputs", "non_cond_frame_outputs"):
            st["output_dict"][k].clear()
            st["consolidated_frame_inds"][k].clear()
        st["tracking_has_started"] = False
        st["frames_already_tracked"].clear()

    def _clear_non_cond_mem_around_input(self, st, frame_idx):
        """Drop non-conditioning memories within the memory window of an edited frame
        (sam2_video_predictor.py:1152-1172)."""
        r = self.memory_temporal_stride_for_eval
        lo, hi = frame_idx - r * self.num_maskmem, frame_idx + r * self.num_maskmem
        for t in range(lo, hi + 1):
            st["output_dict"]["non_cond_frame_outputs"].pop(t, None)
            for per in st["output_dict_per_obj"].values():
                per["non_cond_frame_outputs"].pop(t, None)

    # ------------------------------------------------------------------ per-frame work
    def _get_image_feature(self, st, frame_idx, batch_size):
        """Backbone features of a frame, expanded over objects (sam2_video_predictor.py:879-910)."""
        image, backbone_out = st["cached_features"].get(frame_idx, (None, None))
        if backbone_out is None:
            dev = st["device"]
            if st["feature_source"] is not None:
                image, backbone_out = None, st["feature_source"].frame_features(frame_idx, dev)
            else:
                image = st["images"][frame_idx].to(dev).float().unsqueeze(0)
                backbone_out = self.forward_image(image)
            st["cached_features"] = {frame_idx: (image, backbone_out)}
        expanded = {
            "backbone_fpn": [f.expand(batch_size, -1, -1, -1) for f in backbone_out["backbone_fpn"]],
            "vision_pos_enc": [p.expand(batch_size, -1, -1, -1) for p in backbone_out["vision_pos_enc"]],
        }
        feats = self._prepare_backbone_features(expanded)
        image = None if image is None else image.expand(batch_size, -1, -1, -1)
        return (image,) + feats

    def _run_single_frame_inference(self, inference_state, output_dict, frame_idx, batch_size, is_init_cond_frame,
                                    point_inputs, mask_inputs, reverse, run_mem_encoder, prev_sam_mask_logits=None):
        """One tracked frame -> compact state entry (sam2_video_predictor.py:912-978)."""
        st = inference_state
        _, _, feats, pos, sizes = self._get_image_feature(st, frame_idx, batch_size)
        assert point_inputs is None or mask_inputs is None
        cur = self.track_step(frame_idx=frame_idx, is_init_cond_frame=is_init_cond_frame, current_vision_feats=feats,
                              current_vision_pos_embeds=pos, feat_sizes=sizes, point_inputs=point_inputs,
                              mask_inputs=mask_inputs, output_dict=output_dict, num_frames=st["num_frames"],
                              track_in_reverse=reverse, run_mem_encoder=run_mem_encoder,
                              prev_sam_mask_logits=prev_sam_mask_logits)
        store = st["storage_device"]
        mem, rows = cur["maskmem_features"], cur["maskmem_rows"]
        if mem is not None:
            mem = mem.to(torch.bfloat16).to(store, non_blocking=True)       # memories are kept in bf16 (:956)
            rows = None if rows is None else rows.to(store, non_blocking=True)
        pred_gpu = cur["pred_masks"]
        if self.fill_hole_area > 0:
            pred_gpu = fill_holes_in_mask_scores(pred_gpu, self.fill_hole_area)  # raises on failure, never skips
        compact = {
            "maskmem_features": mem, "maskmem_rows": rows,
            "maskmem_pos_enc": self._get_maskmem_pos_enc(st, cur),
            "pred_masks": pred_gpu.to(store, non_blocking=True),
            "obj_ptr": cur["obj_ptr"], "object_score_logits": cur["object_score_logits"],
        }
        return compact, pred_gpu

    def _run_memory_encoder(self, st, frame_idx, batch_size, high_res_masks, object_score_logits, is_mask_from_pts,
                            low_res_masks=None):
        """Re-encode a consolidated prompt frame (sam2_video_predictor.py:980-1014)."""
        _, _, feats, _, sizes = self._get_image_feature(st, frame_idx, batch_size)
        if low_res_masks is not None:
            mem, rows, pos = self._encode_new_memory_low_res(feats, low_res_masks, object_score_logits, is_mask_from_pts)
        else:
            mem, pos = self._encode_new_memory(feats, sizes, high_res_masks, object_score_logits, is_mask_from_pts)
            rows = None
        store = st["storage_device"]
        mem = mem.to(torch.bfloat16).to(store, non_blocking=True)
        rows = None if rows is None else rows.to(store, non_blocking=True)
        return mem, rows, self._get_maskmem_pos_enc(st, {"maskmem_pos_enc": pos})

    def _get_maskmem_pos_enc(self, st, cur):
        """The positional encoding of a memory is a constant: keep one copy per session (:1016-1039)."""
        pos = cur["maskmem_pos_enc"]
        if pos is None:
            return None
        consts = st["constants"]
        if "maskmem_pos_enc" not in consts:
            consts["maskmem_pos_enc"] = [x[0:1].clone() for x in pos]
        return [x.expand(pos[0].size(0), -1, -1, -1) for x in consts["maskmem_pos_enc"]]
