// B200Renderer: the reference's `Renderer` (ts/src/lib/renderer.ts:4-8) on libswfr_b200.so.
//
// Mirrors CanvasRenderer (ts/src/lib/renderers/canvas-renderer.ts:48-145): definitions are compiled the first time they
// are drawn and cached per tag object; render(stage) walks nothing here - the library flattens the display tree exactly
// as renderStage / drawContainer / drawShape / drawMorphShape do (depth first, CTM' = CTM x M with save / restore).
import { JsonValueWriter } from "kryo/writers/json-value";
import { Matrix } from "swf-tree/matrix";
import { $DefineMorphShape, $DefineShape, DefineBitmap, DefineMorphShape, DefineShape } from "swf-tree/tags";
import { DisplayObject } from "../display/display-object";
import { DisplayObjectType } from "../display/display-object-type";
import { Stage } from "../display/stage";
import { Renderer } from "../renderer";

// tslint:disable-next-line:no-var-requires
const native: { Native: NativeCtor } = require("../../../build/Release/swfr_b200.node");

interface NativeRenderer {
  registerShape(tagJson: object, morph: boolean): number;
  addBitmap(id: number, data: Uint8Array): void; // image/x-swf-bmp bytes
  render(width: number, height: number, background: Uint8Array | undefined, nodes: Int32Array, rootCount: number): void;
  readImage(premultiplied: boolean): Uint8ClampedArray; // width * height * 4, RGBA
  toPng(): Uint8Array; // what canvas.toBuffer("image/png") of the reference test holds
  toPam(): Uint8Array; // imageDataToPam (ts/src/lib/image-data-to-pam.ts:8-30)
  close(): void;
}
type NativeCtor = new (width: number, height: number, device: number) => NativeRenderer;

const JSON_VALUE_WRITER: JsonValueWriter = new JsonValueWriter();

/** Integers per display-tree node in the packed form (see packStage). */
const NODE_INTS: number = 12;

export class B200Renderer implements Renderer {
  public readonly width: number;
  public readonly height: number;
  private readonly native: NativeRenderer;
  private readonly shapeIds: WeakMap<DefineShape, number>;
  private readonly morphShapeIds: WeakMap<DefineMorphShape, number>;

  constructor(width: number, height: number, device: number = 0) {
    this.width = width;
    this.height = height;
    this.native = new native.Native(width, height, device);
    this.shapeIds = new WeakMap();
    this.morphShapeIds = new WeakMap();
  }

  render(stage: Stage): void {
    if (stage.width !== this.width || stage.height !== this.height) {
      throw new Error("StageSizeMismatch");
    }
    const bg: Uint8Array | undefined = stage.backgroundColor === undefined
      ? undefined
      : Uint8Array.of(stage.backgroundColor.r, stage.backgroundColor.g, stage.backgroundColor.b, stage.backgroundColor.a);
    this.native.render(stage.width, stage.height, bg, this.packStage(stage), stage.children.length);
  }

  async addBitmap(tag: DefineBitmap): Promise<void> {
    // node-canvas-bitmap-service.ts:14-37: only image/x-swf-bmp is implemented by the reference
    if (tag.mediaType !== "image/x-swf-bmp") {
      throw new Error("NotImplementedBitmapType");
    }
    this.native.addBitmap(tag.id, tag.data);
  }

  /** Straight-alpha RGBA8 pixels of the last render (CanvasRenderingContext2D.getImageData semantics). */
  getImageData(): { width: number; height: number; data: Uint8ClampedArray } {
    return {width: this.width, height: this.height, data: this.native.readImage(false)};
  }

  toPng(): Uint8Array {
    return this.native.toPng();
  }

  toPam(): Uint8Array {
    return this.native.toPam();
  }

  close(): void {
    this.native.close();
  }

  private shapeId(tag: DefineShape): number {
    let id: number | undefined = this.shapeIds.get(tag);
    if (id === undefined) {
      id = this.native.registerShape($DefineShape.write(JSON_VALUE_WRITER, tag), false);
      this.shapeIds.set(tag, id);
    }
    return id;
  }

  private morphShapeId(tag: DefineMorphShape): number {
    let id: number | undefined = this.morphShapeIds.get(tag);
    if (id === undefined) {
      id = this.native.registerShape($DefineMorphShape.write(JSON_VALUE_WRITER, tag), true);
      this.morphShapeIds.set(tag, id);
    }
    return id;
  }

  /**
   * The display tree as NODE_INTS integers per node, breadth first so that the children of a node are contiguous:
   *   [type, id, hasMatrix, scaleX, scaleY, rotateSkew0, rotateSkew1, translateX, translateY, ratio (float32 bits),
   *    childCount, firstChild]
   * `type` is the reference's DisplayObjectType value (Container 0, MorphShape 1, Shape 2 = swfr_display_object_type);
   * matrix entries are the swf-tree integers (Sfixed16P16 epsilons, twips).  The first stage.children.length nodes are
   * the stage's children.
   */
  private packStage(stage: Stage): Int32Array {
    const queue: ReadonlyArray<DisplayObject>[] = [stage.children];
    const owners: number[] = [-1];
    const out: number[] = [];
    const f32: Float32Array = new Float32Array(1);
    const bits: Int32Array = new Int32Array(f32.buffer);
    for (let q: number = 0; q < queue.length; q++) {
      const first: number = out.length / NODE_INTS;
      if (owners[q] >= 0) {
        out[owners[q] * NODE_INTS + 11] = first;
      }
      for (const node of queue[q]) {
        const m: Matrix | undefined = node.matrix;
        let id: number = 0;
        let childCount: number = 0;
        f32[0] = 0;
        switch (node.type) {
          case DisplayObjectType.Container:
            childCount = node.children.length;
            queue.push(node.children);
            owners.push(out.length / NODE_INTS);
            break;
          case DisplayObjectType.MorphShape:
            id = this.morphShapeId(node.definition);
            f32[0] = node.ratio;
            break;
          case DisplayObjectType.Shape:
            id = this.shapeId(node.definition);
            break;
          default:
            throw new Error("UnexpectedDisplayObjectType"); // canvas-renderer.ts:91-92
        }
        out.push(
          node.type, id, m === undefined ? 0 : 1,
          m === undefined ? 65536 : m.scaleX.epsilons, m === undefined ? 65536 : m.scaleY.epsilons,
          m === undefined ? 0 : m.rotateSkew0.epsilons, m === undefined ? 0 : m.rotateSkew1.epsilons,
          m === undefined ? 0 : m.translateX, m === undefined ? 0 : m.translateY,
          bits[0], childCount, 0,
        );
      }
    }
    return Int32Array.from(out);
  }
}
