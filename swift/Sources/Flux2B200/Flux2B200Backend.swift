// Flux2B200Backend.swift — the Swift shim a maintainer adds to Flux2Core so that the model classes keep their public
// signatures (Flux2Transformer2DModel.callAsFunction, FlowMatchEulerScheduler.step, AutoencoderKLFlux2.decode,
// Flux2StepHook) and forward to the C ABI of include/flux2b.h. It cannot be compiled in the build container (no
// swiftc, no mlx-swift); it is written against mlx-swift 0.31.6's MLXArray API (asArray / asData / init(data:shape:)).
//
// Wiring (Package.swift of the reference):
//   .target(name: "CFlux2B", path: "Sources/CFlux2B", publicHeadersPath: "include",
//           linkerSettings: [.unsafeFlags(["-L/path/to/flux2b/lib", "-lflux2b"])]),
//   .target(name: "Flux2Core", dependencies: [..., "CFlux2B"])
import CFlux2B
import Foundation
import MLX

public enum Flux2B200Error: Error { case status(Int32, String) }

@inline(__always) func f2bCheck(_ rc: Int32) throws {
    // flux2b_status maps 1:1 onto Flux2Error (Flux2Core.swift:14-40)
    guard rc >= 0 else {
        let msg = String(cString: flux2b_last_error())
        switch rc {
        case -1: throw Flux2Error.modelNotLoaded(msg)
        case -2: throw Flux2Error.invalidConfiguration(msg)
        case -3: throw Flux2Error.insufficientMemory(required: 0, available: 0)
        case -4: throw Flux2Error.weightLoadingFailed(msg)
        case -5: throw Flux2Error.imageProcessingFailed(msg)
        case -7: throw Flux2Error.generationCancelled
        default: throw Flux2Error.generationFailed(msg)
        }
    }
}

/// One flux2b context == one GPU == the transformer + VAE of one Flux2Pipeline.
public final class Flux2B200Context: @unchecked Sendable {
    let handle: OpaquePointer

    public init(device: Int32 = 0, transformer: Flux2TransformerConfig?, vae: VAEConfig?, quantization: TransformerQuantization) throws {
        var dit = flux2b_dit_config()
        if let t = transformer {
            dit.patch_size = Int32(t.patchSize); dit.in_channels = Int32(t.inChannels); dit.out_channels = Int32(t.outChannels)
            dit.num_layers = Int32(t.numLayers); dit.num_single_layers = Int32(t.numSingleLayers)
            dit.attention_head_dim = Int32(t.attentionHeadDim); dit.num_attention_heads = Int32(t.numAttentionHeads)
            dit.joint_attention_dim = Int32(t.jointAttentionDim); dit.guidance_embeds = t.guidanceEmbeds ? 1 : 0
            dit.axes_dims_rope = (Int32(t.axesDimsRope[0]), Int32(t.axesDimsRope[1]), Int32(t.axesDimsRope[2]), Int32(t.axesDimsRope[3]))
            dit.rope_theta = t.ropeTheta; dit.mlp_ratio = t.mlpRatio
        }
        var v = flux2b_vae_config()
        if let c = vae {
            v.in_channels = Int32(c.inChannels); v.out_channels = Int32(c.outChannels); v.latent_channels = Int32(c.latentChannels)
            v.layers_per_block = Int32(c.layersPerBlock); v.norm_num_groups = Int32(c.normNumGroups); v.norm_eps = c.normEps
            let ch = c.effectiveDecoderChannels
            v.decoder_channels = (Int32(ch[0]), Int32(ch[1]), Int32(ch[2]), Int32(ch[3]))
            let ec = c.blockOutChannels
            v.encoder_channels = (Int32(ec[0]), Int32(ec[1]), Int32(ec[2]), Int32(ec[3]))
        }
        let q: Int32 = { switch quantization { case .bf16: return 0; case .qint8: return 1; case .int4: return 2
                                                case .mxfp8: return 3; case .mxfp4: return 4; case .nvfp4: return 5 } }()
        var h: OpaquePointer?
        try withUnsafePointer(to: &dit) { dp in try withUnsafePointer(to: &v) { vp in
            try f2bCheck(flux2b_create(device, transformer != nil ? dp : nil, vae != nil ? vp : nil, q, &h)) } }
        handle = h!
    }
    deinit { flux2b_destroy(handle) }

    /// Flux2WeightLoader.applyTransformerWeights / applyVAEWeights (WeightLoader.swift:567-623): same flattened keys.
    public func setWeights(_ weights: [String: MLXArray]) throws {
        for (key, w) in weights {
            eval(w)
            let code: Int32 = { switch w.dtype { case .float32: return 0; case .float16: return 1; case .bfloat16: return 2
                                                   case .uint32: return 3; case .uint8: return 4; default: return 5 } }()
            var shape = w.shape.map { Int64($0) }
            let data = w.asData(access: .noCopyIfContiguous)
            try data.data.withUnsafeBytes { raw in
                try f2bCheck(flux2b_set_tensor(handle, key, raw.baseAddress, code, &shape, Int32(shape.count)))
            }
        }
    }
    public func finalize() throws { try f2bCheck(flux2b_finalize_weights(handle)) }
}

extension Flux2Transformer2DModel {
    /// Drop-in body for callAsFunction (Flux2Transformer.swift:123-327) when a B200 context is attached.
    public func b200Forward(_ ctx: Flux2B200Context, hiddenStates: MLXArray, encoderHiddenStates: MLXArray, timestep: MLXArray,
                            guidance: MLXArray?, imgIds: MLXArray, txtIds: MLXArray) throws -> MLXArray {
        let B = hiddenStates.dim(0), sImg = hiddenStates.dim(1), sTxt = encoderHiddenStates.dim(1)
        let hid = hiddenStates.asType(.float32).asArray(Float.self)
        let enc = encoderHiddenStates.asType(.float32).asArray(Float.self)
        let t = timestep.asType(.float32).asArray(Float.self)
        let g = guidance?.asType(.float32).asArray(Float.self)
        let ii = imgIds.asType(.int32).asArray(Int32.self), ti = txtIds.asType(.int32).asArray(Int32.self)
        var out = [Float](repeating: 0, count: B * sImg * config.outChannels)
        try f2bCheck(flux2b_dit_forward(ctx.handle, Int32(B), Int32(sImg), Int32(sTxt), hid, enc, 0, t, g, ii, ti, &out))
        return MLXArray(out, [B, sImg, config.outChannels])
    }
}

extension Flux2Pipeline {
    /// Replaces the T2I / I2I / KV loop bodies (Flux2Pipeline.swift:1933-2052, 1696-1808, 1555-1683): the whole loop runs on the
    /// device; control returns to Swift only when a Flux2StepHook is installed. `kvCache` = `model.supportsKVCache` (:1555).
    func b200Denoise(_ ctx: Flux2B200Context, latents: MLXArray, textEmbeddings: MLXArray, negativeEmbeddings: MLXArray?,
                     sigmas: [Float], guidance: Float?, cfgScale: Float, height: Int, width: Int,
                     refLatents: MLXArray?, refIds: MLXArray?, kvCache: Bool = false, onStep: Flux2StepHook?) throws -> MLXArray {
        var x = latents.asType(.float32).asArray(Float.self)
        let enc = textEmbeddings.asType(.float32).asArray(Float.self)
        let neg = negativeEmbeddings?.asType(.float32).asArray(Float.self)
        let refs = refLatents?.asType(.float32).asArray(Float.self)
        let rids = refIds?.asType(.int32).asArray(Int32.self)
        var gval = guidance ?? 0
        let seq = latents.dim(1)
        // C callback -> Swift closure: latents arrive as a host float buffer, whatever is left in it replaces them
        final class Box { let hook: Flux2StepHook; let seq: Int; init(_ h: @escaping Flux2StepHook, _ s: Int) { hook = h; seq = s } }
        let box = onStep.map { Box($0, seq) }
        let cHook: flux2b_step_hook? = box == nil ? nil : { sc, lat, n, user in
            let b = Unmanaged<Box>.fromOpaque(user!).takeUnretainedValue()
            let c = sc!.pointee
            let ctx = Flux2StepContext(stepIdx: Int(c.step_idx), totalSteps: Int(c.total_steps), sigma: c.sigma,
                                       sigmaNext: c.sigma_next, height: Int(c.height), width: Int(c.width), isI2I: c.is_i2i != 0)
            let inArr = MLXArray(UnsafeBufferPointer(start: lat, count: n), [1, b.seq, n / b.seq])
            let outArr = b.hook(ctx, inArr).asType(.float32)
            eval(outArr)
            outArr.asArray(Float.self).withUnsafeBufferPointer { lat!.update(from: $0.baseAddress!, count: n) }
            return 0
        }
        try sigmas.withUnsafeBufferPointer { sp in
            var p = flux2b_denoise_params()
            p.height = Int32(height); p.width = Int32(width); p.num_sigmas = Int32(sigmas.count); p.sigmas = sp.baseAddress
            p.cfg_scale = cfgScale; p.enc_dtype = 0; p.S_txt = Int32(textEmbeddings.dim(1)); p.S_ref = Int32(refLatents?.dim(1) ?? 0)
            p.hook = cHook
            p.kv_cache = kvCache ? 1 : 0
            p.S_txt_uncond = Int32(negativeEmbeddings?.dim(1) ?? 0)   // uncondTextIds follow the negative prompt's own length (:1690)
            p.hook_user = box.map { UnsafeMutableRawPointer(Unmanaged.passUnretained($0).toOpaque()) }
            try enc.withUnsafeBytes { e in
                p.enc = e.baseAddress
                try withExtendedLifetime((neg, refs, rids)) {
                    neg?.withUnsafeBytes { p.enc_uncond = $0.baseAddress }
                    refs?.withUnsafeBufferPointer { p.ref_latents = $0.baseAddress }
                    rids?.withUnsafeBufferPointer { p.ref_ids = $0.baseAddress }
                    try withUnsafePointer(to: &gval) { gp in
                        p.guidance = guidance != nil ? gp : nil
                        try f2bCheck(flux2b_denoise(ctx.handle, &p, &x))
                    }
                }
            }
        }
        return MLXArray(x, latents.shape)
    }
}

extension AutoencoderKLFlux2 {
    /// Drop-in body for decode (AutoencoderKL.swift:129-143).
    public func b200Decode(_ ctx: Flux2B200Context, _ z: MLXArray) throws -> MLXArray {
        let B = z.dim(0), h = z.dim(2), w = z.dim(3)
        let lat = z.asType(.float32).asArray(Float.self)
        var img = [Float](repeating: 0, count: B * 3 * 64 * h * w)
        try f2bCheck(flux2b_vae_decode(ctx.handle, Int32(B), Int32(h), Int32(w), lat, &img))
        return MLXArray(img, [B, 3, 8 * h, 8 * w])
    }

    /// Body of `AutoencoderKLFlux2.encode(_:samplePosterior:)` (VAE/AutoencoderKL.swift:90). `noise` = nil is
    /// `samplePosterior: false`; otherwise the caller draws `MLXRandom.normal(mean.shape)` and hands it over.
    public func b200Encode(_ ctx: Flux2B200Context, _ x: MLXArray, noise: MLXArray? = nil) throws -> MLXArray {
        let B = x.dim(0), H = x.dim(2), W = x.dim(3)
        let img = x.asType(.float32).asArray(Float.self)
        let nz = noise?.asType(.float32).asArray(Float.self)
        var lat = [Float](repeating: 0, count: B * 32 * (H / 8) * (W / 8))
        try f2bCheck(flux2b_vae_encode(ctx.handle, Int32(B), Int32(H), Int32(W), img, nz, &lat))
        return MLXArray(lat, [B, 32, H / 8, W / 8])
    }

    /// `Flux2Pipeline.encodeImageToPackedSequence` after `preprocessImageForVAE` (Flux2Pipeline+ChainHelpers.swift:75-101).
    public func b200EncodeToPackedSequence(_ ctx: Flux2B200Context, _ x: MLXArray) throws -> MLXArray {
        let B = x.dim(0), H = x.dim(2), W = x.dim(3)
        let img = x.asType(.float32).asArray(Float.self)
        var seq = [Float](repeating: 0, count: B * (H / 16) * (W / 16) * 128)
        try f2bCheck(flux2b_encode_image_to_sequence(ctx.handle, Int32(B), Int32(H), Int32(W), img, nil, &seq))
        return MLXArray(seq, [B, (H / 16) * (W / 16), 128])
    }
}

// MARK: - Text-embedding producer (SURVEY §8 f-4). Lives in the FluxTextEncoders target in practice (it only needs CFlux2B).

/// One flux2b text-encoder context == the decoder layers of Qwen3Model / MistralModel that the embedding extractors run.
public final class Flux2B200TextEncoder: @unchecked Sendable {
    let handle: OpaquePointer
    public let hiddenSize: Int

    /// `bits` / `groupSize` come from the checkpoint's config.json "quantization" block (Qwen3Model.swift:311-345): mlx-community
    /// 8-bit / 4-bit (group 64, affine) map to FLUX2B_QINT8 / FLUX2B_INT4; an unquantized checkpoint is FLUX2B_BF16.
    public init(device: Int32 = 0, config c: Qwen3TextConfig, quantBits: Int? = nil) throws {
        var t = flux2b_te_config()
        t.vocab_size = Int32(c.vocabSize); t.hidden_size = Int32(c.hiddenSize); t.intermediate_size = Int32(c.intermediateSize)
        t.num_layers = Int32(c.numHiddenLayers); t.num_heads = Int32(c.numAttentionHeads); t.num_kv_heads = Int32(c.numKeyValueHeads)
        t.head_dim = Int32(c.headDim); t.qk_norm = 1; t.rms_norm_eps = c.rmsNormEps; t.rope_theta = c.ropeTheta
        t.max_position_embeddings = 0
        let q: Int32 = quantBits == 8 ? 1 : (quantBits == 4 ? 2 : 0)
        var h: OpaquePointer?
        try withUnsafePointer(to: &t) { tp in try f2bCheck(flux2b_te_create(device, tp, q, &h)) }
        handle = h!
        hiddenSize = c.hiddenSize
    }
    /// Flux.2 Dev: Mistral Small 3.2 as EmbeddingExtractor.extractFluxEmbeddings uses it (EmbeddingExtractor.swift:202-285; LEFT
    /// padding, hidden states 10 / 20 / 30). No QK-norm; the Llama-4 query scale (MistralAttention.swift:15-32) is 1 below
    /// originalMaxPositionEmbeddings, which the library checks instead of computing.
    public init(device: Int32 = 0, mistral c: MistralTextConfig, quantBits: Int? = nil) throws {
        var t = flux2b_te_config()
        t.vocab_size = Int32(c.vocabSize); t.hidden_size = Int32(c.hiddenSize); t.intermediate_size = Int32(c.intermediateSize)
        t.num_layers = Int32(c.numHiddenLayers); t.num_heads = Int32(c.numAttentionHeads); t.num_kv_heads = Int32(c.numKeyValueHeads)
        t.head_dim = Int32(c.headDim); t.qk_norm = 0; t.rms_norm_eps = c.rmsNormEps; t.rope_theta = c.ropeTheta
        t.max_position_embeddings = Int32(c.originalMaxPositionEmbeddings)
        let q: Int32 = quantBits == 8 ? 1 : (quantBits == 4 ? 2 : 0)
        var h: OpaquePointer?
        try withUnsafePointer(to: &t) { tp in try f2bCheck(flux2b_te_create(device, tp, q, &h)) }
        handle = h!
        hiddenSize = c.hiddenSize
    }
    deinit { flux2b_destroy(handle) }

    /// Same keys as the checkpoint / Module paths ("model.layers.3.self_attn.q_proj.weight", ".scales", ".biases", ...).
    /// Layers beyond the deepest extracted hidden state (27 for Klein) may be skipped.
    public func setWeights(_ weights: [String: MLXArray]) throws {
        for (key, w) in weights {
            eval(w)
            let code: Int32 = { switch w.dtype { case .float32: return 0; case .float16: return 1; case .bfloat16: return 2
                                                   case .uint32: return 3; case .uint8: return 4; default: return 5 } }()
            var shape = w.shape.map { Int64($0) }
            let data = w.asData(access: .noCopyIfContiguous)
            try data.data.withUnsafeBytes { raw in
                try f2bCheck(flux2b_set_tensor(handle, key, raw.baseAddress, code, &shape, Int32(shape.count)))
            }
        }
    }
    public func finalize() throws { try f2bCheck(flux2b_finalize_weights(handle)) }

    /// Drop-in body for Qwen3Model.forwardWithHiddenStates(_:layerIndices:attentionMask:) (Qwen3Model.swift:104-191) followed by
    /// the concatenation of KleinEmbeddingExtractor.extractKleinEmbeddings step 9-10 (KleinEmbeddingExtractor.swift:111-121).
    /// Tokenisation, chat template, truncation and RIGHT padding (steps 1-7) stay in the extractor.
    public func hiddenStates(inputIds: MLXArray, layerIndices: [Int], attentionMask: MLXArray?) throws -> MLXArray {
        let B = inputIds.dim(0), S = inputIds.dim(1)
        let ids = inputIds.asType(.int32).asArray(Int32.self)
        let mask = attentionMask?.asType(.int32).asArray(Int32.self)
        let layers = layerIndices.map { Int32($0) }
        var out = [Float](repeating: 0, count: B * S * layers.count * hiddenSize)
        try f2bCheck(flux2b_te_hidden_states(handle, Int32(B), Int32(S), ids, mask, layers, Int32(layers.count), &out, 0))
        return MLXArray(out, [B, S, layers.count * hiddenSize])
    }
}
